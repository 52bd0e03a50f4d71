#include "slab_pipeline.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sys/mman.h>
#include <unistd.h>
#include <zlib.h>

#include "../../include/tcb200.h"

namespace tc {

namespace {

typedef std::chrono::steady_clock Clock;
double seconds(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

int typeOf(Encoding e) { return e == Encoding::BC1 ? 1 : e == Encoding::BC2 ? 2 : e == Encoding::BC3 ? 3 : 0; }

// A fixed set of worker threads; run(n, fn) executes fn(0..n-1) on them and returns when all are done.  Two threads
// (feeder and harvester) may call run() at the same time.
class Workers
{
public:
  explicit Workers(int n)
  {
    for (int i = 0; i < n; ++i) m_threads.emplace_back([this] { loop(); });
  }
  ~Workers()
  {
    { std::lock_guard<std::mutex> lock(m_mutex); m_stop = true; }
    m_wake.notify_all();
    for (std::thread& t : m_threads) t.join();
  }
  void run(int n, const std::function<void(int)>& fn)
  {
    if (n <= 0) return;
    Batch batch;
    batch.fn = &fn; batch.left = n;
    {
      std::lock_guard<std::mutex> lock(m_mutex);
      for (int i = 0; i < n; ++i) m_tasks.push_back(Task{ &batch, i });
    }
    m_wake.notify_all();
    std::unique_lock<std::mutex> lock(m_mutex);
    batch.done.wait(lock, [&] { return batch.left == 0; });
  }

private:
  struct Batch { const std::function<void(int)>* fn; int left; std::condition_variable done; };
  struct Task { Batch* batch; int index; };
  void loop()
  {
    std::unique_lock<std::mutex> lock(m_mutex);
    for (;;) {
      m_wake.wait(lock, [&] { return m_stop || !m_tasks.empty(); });
      if (m_tasks.empty()) { if (m_stop) return; continue; }
      Task t = m_tasks.front();
      m_tasks.erase(m_tasks.begin());
      lock.unlock();
      (*t.batch->fn)(t.index);
      lock.lock();
      if (--t.batch->left == 0) t.batch->done.notify_all();
    }
  }
  std::mutex m_mutex;
  std::condition_variable m_wake;
  std::vector<Task> m_tasks;
  std::vector<std::thread> m_threads;
  bool m_stop = false;
};

// Level-9 one-shot zlib stream per tile, byte-identical to Compression::deflate (tileconv/compress.cpp:36,54-67),
// with the stream state reused across tiles (deflateReset) instead of rebuilt.
class RecordDeflater
{
public:
  RecordDeflater() { std::memset(&m_z, 0, sizeof(m_z)); m_ok = deflateInit(&m_z, 9) == Z_OK; }
  ~RecordDeflater() { if (m_ok) deflateEnd(&m_z); }
  uint32_t run(const uint8_t* src, uint32_t srcSize, uint8_t* dst, uint32_t dstSize)
  {
    if (!m_ok) return 0;
    m_z.avail_in = srcSize; m_z.next_in = const_cast<Bytef*>(src);
    m_z.avail_out = dstSize; m_z.next_out = dst;
    const int rc = ::deflate(&m_z, Z_FINISH);
    const uint32_t written = dstSize - m_z.avail_out;
    deflateReset(&m_z);
    return rc == Z_STREAM_END ? written : 0;
  }
private:
  z_stream m_z;
  bool m_ok;
};

void wr32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

bool pwriteAll(int fd, const uint8_t* p, size_t n, unsigned long long off)
{
  while (n > 0) {
    const ssize_t w = ::pwrite(fd, p, n, (off_t)off);
    if (w <= 0) return false;
    p += w; n -= (size_t)w; off += (unsigned long long)w;
  }
  return true;
}

}  // namespace

bool slabPipelineApplies(const Options& options) noexcept
{
  if (typeOf(options.getEncoding()) == 0) return false;
  const char* env = std::getenv("TCB200_TILEDATA_ROUTE");
  return !(env && env[0] == '1');
}

bool runSlabPipeline(const Options& options, TileFeed& feed, int fd, unsigned long long offset, std::string& error,
                     PipelineTimes* times) noexcept
{
  try {
    const Clock::time_point tStart = Clock::now();
    const int type = typeOf(options.getEncoding()), quality = options.getEncodingQuality();
    const unsigned count = feed.count();
    const bool deflate = options.isDeflate();
    if (type == 0) { error = "slab pipeline: not a BC1/BC2/BC3 encode"; return false; }

    int devices = tcb200_device_count();
    if (const char* env = std::getenv("TCB200_GPUS")) { const int v = std::atoi(env); if (v > 0 && v < devices) devices = v; }
    if (devices <= 0) { error = "no usable CUDA device (this encoder has no CPU fallback)"; return false; }
    const int threads = std::max(1, options.getThreads());

    // slab size: large enough to amortise per-slab costs at RangeFit rates, small enough that every device gets
    // several slabs of a small file
    int slabTiles = (int)std::min<unsigned>(32768u, std::max<unsigned>(256u, (count + (unsigned)devices * 4u - 1u) / ((unsigned)devices * 4u)));
    if (const char* env = std::getenv("TCB200_SLAB_TILES")) { const int v = std::atoi(env); if (v > 0) slabTiles = v; }
    const unsigned jobs = (count + (unsigned)slabTiles - 1u) / (unsigned)slabTiles;
    if ((unsigned)devices > jobs) devices = (int)std::max(1u, jobs);

    struct Device { tcb200_ctx* ctx = nullptr; int slabs = 0; };
    std::vector<Device> dev((size_t)devices);
    struct Cleanup { std::vector<Device>& d; ~Cleanup() { for (Device& g : d) tcb200_destroy(g.ctx); } } cleanup{ dev };
    for (int g = 0; g < devices; ++g) {
      dev[(size_t)g].ctx = tcb200_create(g, type, quality, slabTiles);
      if (!dev[(size_t)g].ctx) { error = std::string("tcb200_create: ") + tcb200_last_error(nullptr); return false; }
      dev[(size_t)g].slabs = tcb200_slab_count(dev[(size_t)g].ctx);
      for (int s = 0; s < dev[(size_t)g].slabs; ++s)       // allocate the pinned slabs up front (outside the timed stages)
        if (!tcb200_slab_input(dev[(size_t)g].ctx, s)) { error = std::string("pinned slab: ") + tcb200_last_error(dev[(size_t)g].ctx); return false; }
    }
    const int stride = tcb200_record_stride(dev[0].ctx);
    const int slabsPerDev = dev[0].slabs;
    const unsigned ring = (unsigned)devices * (unsigned)slabsPerDev;     // slab jobs that can be in flight

    Workers workers(threads);
    // per ring slot: staging buffer for the framed tiles of one slab (u32 size + payload each; 2x for zlib's bound)
    const size_t frameCap = (size_t)slabTiles * (4u + (size_t)stride * (deflate ? 2u : 1u));
    std::vector<std::vector<uint8_t>> frame(2);           // the harvester frames one slab while it writes the previous
    for (auto& f : frame) f.resize(frameCap);

    // -u: every framed tile has a known size (4 + record_size(w,h)), so the output file is sized up front and the
    // worker threads frame straight into a shared mapping of it -- buffered write()s to one file are serialised
    // by the inode lock, page faults on a mapping are not.  With zlib the sizes are only known after compression
    // and the framed slabs are written with pwrite (deflate is the bottleneck there).
    uint8_t* map = nullptr;
    size_t mapBytes = 0;
    unsigned long long mapBase = 0;          // file offset of map[0] (page aligned)
    struct Unmap { uint8_t*& p; size_t& n; ~Unmap() { if (p) ::munmap(p, n); } } unmap{ map, mapBytes };
    if (!deflate) {
      unsigned long long payload = 0;
      for (unsigned i = 0; i < count; ++i) {
        uint16_t w = 64, h = 64;
        feed.dims(i, w, h);
        payload += 4u + (unsigned long long)tcb200_record_size(type, w, h);
      }
      if (::ftruncate(fd, (off_t)(offset + payload)) == 0) {
        const unsigned long long page = (unsigned long long)::sysconf(_SC_PAGESIZE);
        mapBase = offset / page * page;
        mapBytes = (size_t)(offset + payload - mapBase);
        void* p = ::mmap(nullptr, mapBytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)mapBase);
        map = (p == MAP_FAILED) ? nullptr : static_cast<uint8_t*>(p);      // falls back to pwrite when mapping fails
      }
    }

    PipelineTimes t;
    t.devices = devices; t.slabTiles = slabTiles; t.threads = threads; t.tiles = count;
    t.setup = seconds(tStart, Clock::now());
    const Clock::time_point tRun = Clock::now();

    std::mutex m;
    std::condition_variable cvSubmitted, cvFree;
    unsigned submitted = 0, harvested = 0;
    std::atomic<bool> failed(false);
    std::string feederError;

    auto jobFirst = [&](unsigned k) { return k * (unsigned)slabTiles; };
    auto jobTiles = [&](unsigned k) { return std::min<unsigned>((unsigned)slabTiles, count - jobFirst(k)); };
    auto jobDev = [&](unsigned k) { return (int)(k % (unsigned)devices); };
    auto jobSlab = [&](unsigned k) { return (int)((k / (unsigned)devices) % (unsigned)slabsPerDev); };

    std::thread feeder([&] {
      for (unsigned k = 0; k < jobs && !failed.load(); ++k) {
        {
          std::unique_lock<std::mutex> lock(m);
          cvFree.wait(lock, [&] { return failed.load() || k < harvested + ring; });
          if (failed.load()) break;
        }
        tcb200_ctx* ctx = dev[(size_t)jobDev(k)].ctx;
        const int slab = jobSlab(k);
        uint8_t* in = tcb200_slab_input(ctx, slab);
        uint16_t* wh = tcb200_slab_sizes(ctx, slab);
        const unsigned first = jobFirst(k), n = jobTiles(k);
        const Clock::time_point t0 = Clock::now();
        const int parts = (int)std::min<unsigned>((unsigned)threads, (n + 255u) / 256u);
        std::atomic<bool> ok(true);
        workers.run(parts, [&](int p) {
          const unsigned lo = (unsigned)((unsigned long long)n * (unsigned)p / (unsigned)parts), hi = (unsigned)((unsigned long long)n * (unsigned)(p + 1) / (unsigned)parts);
          if (!feed.fill(first + lo, hi - lo, in + (size_t)lo * TCB200_TILE_RECORD, wh + (size_t)lo * 2)) ok.store(false);
        });
        const Clock::time_point t1 = Clock::now();
        int rc = ok.load() ? tcb200_slab_submit(ctx, slab, (int)n) : TCB200_E_ARG;
        const Clock::time_point t2 = Clock::now();
        t.fill += seconds(t0, t1); t.submit += seconds(t1, t2);
        if (rc != TCB200_OK) {
          std::lock_guard<std::mutex> lock(m);
          feederError = ok.load() ? std::string("tcb200_slab_submit: ") + tcb200_last_error(ctx) : "Error while reading tile data";
          failed.store(true);
          cvSubmitted.notify_all();
          break;
        }
        { std::lock_guard<std::mutex> lock(m); submitted = k + 1; }
        cvSubmitted.notify_all();
      }
    });

    // ---- harvester: this thread -----------------------------------------------------------------------------
    unsigned long long fileOffset = offset;
    std::vector<uint32_t> sizes((size_t)slabTiles);
    std::vector<unsigned long long> partBytes((size_t)threads + 1);
    for (unsigned k = 0; k < jobs && !failed.load(); ++k) {
      {
        std::unique_lock<std::mutex> lock(m);
        cvSubmitted.wait(lock, [&] { return failed.load() || submitted > k; });
        if (failed.load()) break;
      }
      tcb200_ctx* ctx = dev[(size_t)jobDev(k)].ctx;
      const int slab = jobSlab(k);
      const unsigned n = jobTiles(k);
      const Clock::time_point t0 = Clock::now();
      if (tcb200_slab_wait(ctx, slab) != TCB200_OK) {
        error = std::string("tcb200_slab_wait: ") + tcb200_last_error(ctx);
        failed.store(true); cvFree.notify_all();
        break;
      }
      const Clock::time_point t1 = Clock::now();
      const uint8_t* rec = tcb200_slab_output(ctx, slab);
      const uint16_t* wh = tcb200_slab_sizes(ctx, slab);
      std::vector<uint8_t>& buf = frame[k & 1u];
      const int parts = (int)std::min<unsigned>((unsigned)threads, (n + 127u) / 128u);
      std::atomic<bool> ok(true);
      // each part frames its tiles into its own region of the staging buffer (worst-case spacing), then the regions
      // are written at their final file offsets
      const size_t perTileCap = 4u + (size_t)stride * (deflate ? 2u : 1u);
      // with the output mapped, every part needs its final offset first: a prefix sum over the slab's record sizes
      std::vector<unsigned long long> partStart((size_t)parts + 1, 0);
      if (map) {
        for (int p = 0; p < parts; ++p) {
          const unsigned lo = (unsigned)((unsigned long long)n * (unsigned)p / (unsigned)parts), hi = (unsigned)((unsigned long long)n * (unsigned)(p + 1) / (unsigned)parts);
          unsigned long long bytes = 0;
          for (unsigned i = lo; i < hi; ++i) bytes += 4u + (unsigned long long)tcb200_record_size(type, wh[2 * i], wh[2 * i + 1]);
          partStart[(size_t)p + 1] = partStart[(size_t)p] + bytes;
        }
      }
      workers.run(parts, [&](int p) {
        const unsigned lo = (unsigned)((unsigned long long)n * (unsigned)p / (unsigned)parts), hi = (unsigned)((unsigned long long)n * (unsigned)(p + 1) / (unsigned)parts);
        uint8_t* const dst0 = map ? map + (size_t)(fileOffset + partStart[(size_t)p] - mapBase) : buf.data() + (size_t)lo * perTileCap;
        uint8_t* dst = dst0;
        RecordDeflater z;
        for (unsigned i = lo; i < hi; ++i) {
          const uint32_t recSize = (uint32_t)tcb200_record_size(type, wh[2 * i], wh[2 * i + 1]);
          const uint8_t* src = rec + (size_t)i * (size_t)stride;
          uint32_t payload = recSize;
          if (deflate) {
            payload = z.run(src, recSize, dst + 4, (uint32_t)stride * 2u);
            if (payload == 0) { ok.store(false); return; }
          } else {
            std::memcpy(dst + 4, src, recSize);
          }
          wr32(dst, payload);
          dst += 4u + payload;
        }
        partBytes[(size_t)p] = (unsigned long long)(dst - dst0);
      });
      const Clock::time_point t2 = Clock::now();
      if (!ok.load()) { error = "Error while compressing tile data"; failed.store(true); cvFree.notify_all(); break; }
      // the slab's device buffers and pinned output are free again: let the feeder reuse the slot while we write
      { std::lock_guard<std::mutex> lock(m); harvested = k + 1; }
      cvFree.notify_all();
      std::vector<unsigned long long> partOffset((size_t)parts);
      unsigned long long slabBytes = 0;
      for (int p = 0; p < parts; ++p) { partOffset[(size_t)p] = fileOffset + slabBytes; slabBytes += partBytes[(size_t)p]; }
      if (!map)
        workers.run(parts, [&](int p) {
          const unsigned lo = (unsigned)((unsigned long long)n * (unsigned)p / (unsigned)parts);
          if (!pwriteAll(fd, buf.data() + (size_t)lo * perTileCap, (size_t)partBytes[(size_t)p], partOffset[(size_t)p])) ok.store(false);
        });
      const Clock::time_point t3 = Clock::now();
      if (!ok.load()) { error = "Error while writing tile data"; failed.store(true); cvFree.notify_all(); break; }
      fileOffset += slabBytes;
      t.gpuWait += seconds(t0, t1); t.frame += seconds(t1, t2); t.write += seconds(t2, t3);
      t.bytesIn += (unsigned long long)n * TCB200_TILE_RECORD; t.bytesOut += slabBytes;
    }
    cvFree.notify_all();
    feeder.join();
    if (failed.load()) {
      if (error.empty()) error = feederError.empty() ? "Error while encoding tile data" : feederError;
      return false;
    }
    t.total = seconds(tRun, Clock::now());
    if (times) *times = t;
    return true;
  } catch (const std::exception& e) {
    error = std::string("slab pipeline: ") + e.what();
    return false;
  } catch (...) {
    error = "slab pipeline: unexpected failure";
    return false;
  }
}

}  // namespace tc
