#include "tile_batch_queue.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <memory>

#include <zlib.h>

#include "tcb200.h"

namespace tc {

namespace {

int pixelTypeOf(const Options& options)
{
  switch (options.getEncoding()) {
    case Encoding::BC1: return 1;
    case Encoding::BC2: return 2;
    case Encoding::BC3: return 3;
    default: return 0;
  }
}

// One-shot zlib streams at level 9 into a buffer of 2x the source size, byte-identical to what the
// reference produces per tile (compress.cpp:36,54-67; tiledata.cpp:142-143).  The reference builds a
// fresh Compression object (deflateInit: ~270 KB of state, i.e. an mmap/munmap pair) for every tile; here
// one stream serves a worker's whole share of a slab and is deflateReset between tiles, which yields the
// same bytes without the allocator traffic.
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
    return rc == Z_STREAM_ERROR ? 0 : written;
  }
private:
  z_stream m_z;
  bool m_ok;
};

}  // namespace

TileBatchQueue::TileBatchQueue(unsigned threadNum, unsigned tileNum, int slabTiles, int maxGpus) noexcept
: TileThreadPool(tileNum)
, m_threads(std::max(1u, threadNum))
, m_slabTiles(std::max(1, slabTiles))
, m_maxGpus(maxGpus)
{
  if (const char* env = std::getenv("TCB200_SLAB_TILES")) { int v = std::atoi(env); if (v > 0) m_slabTiles = v; }
  if (const char* env = std::getenv("TCB200_GPUS")) { int v = std::atoi(env); if (v > 0) m_maxGpus = v; }
}

TileBatchQueue::~TileBatchQueue() noexcept
{
  { std::lock_guard<std::mutex> lock(m_inMutex); m_stop.store(true); }
  m_inCv.notify_all(); m_inSpaceCv.notify_all();
  { std::lock_guard<std::mutex> lock(m_flightMutex); }
  m_flightCv.notify_all(); m_slotFreeCv.notify_all();
  { std::lock_guard<std::mutex> lock(m_hostMutex); }
  m_hostWake.notify_all(); m_hostSpaceCv.notify_all();
  if (m_feeder.joinable()) m_feeder.join();
  m_flightCv.notify_all();
  if (m_harvester.joinable()) m_harvester.join();
  for (std::thread& t : m_workers) t.join();
  for (Gpu& g : m_gpus) tcb200_destroy(g.ctx);
}

void TileBatchQueue::setError(const std::string& msg) noexcept
{
  bool first;
  {
    std::lock_guard<std::mutex> lock(m_doneMutex);
    first = m_error.empty();
    m_error = msg;
  }
  // The reference's callers only see a size-0 tile and print TileData::getErrorMsg(), which the pool cannot set
  // for a batch failure (graphics.cpp:127-136): say why on stderr, once per queue, so a CLI user sees the CUDA text.
  if (first && !msg.empty()) std::fprintf(stderr, "tileconv_b200: GPU encode failed: %s\n", msg.c_str());
}

std::string TileBatchQueue::lastError() noexcept
{
  std::lock_guard<std::mutex> lock(m_doneMutex);
  return m_error;
}

int TileBatchQueue::getActiveThreads() noexcept
{
  std::lock_guard<std::mutex> lock(m_doneMutex);
  return m_pending > 0 ? 1 : 0;
}

bool TileBatchQueue::ensureContexts(const Options& options) noexcept
{
  if (!m_gpus.empty()) return true;
  m_type = pixelTypeOf(options);
  m_quality = options.getEncodingQuality();
  m_deflate = options.isDeflate();
  const int visible = tcb200_device_count();
  // m_maxGpus (TCB200_GPUS) = number of encoder contexts.  Normally <= the visible devices; a larger value puts
  // several contexts on one device (round-robin), which only adds slabs in flight -- it is how the multi-context
  // paths of the queue are exercised on a one-GPU box.
  int devices = m_maxGpus > 0 ? std::min(m_maxGpus, 64) : visible;
  if (visible <= 0 || m_type == 0) {
    setError(visible <= 0 ? "no CUDA device (there is no CPU fallback)" : "pixel type is not BC1/BC2/BC3");
    return false;
  }
  for (int d = 0; d < devices; ++d) {
    tcb200_ctx* ctx = tcb200_create(d % visible, m_type, m_quality, m_slabTiles);
    if (ctx == nullptr) {
      setError(tcb200_last_error(nullptr));
      for (Gpu& g : m_gpus) tcb200_destroy(g.ctx);
      m_gpus.clear();
      return false;
    }
    Gpu g; g.ctx = ctx;
    m_gpus.push_back(g);
  }
  const int slabs = tcb200_slab_count(m_gpus[0].ctx);
  for (int s = 0; s < slabs; ++s)
    for (int d = 0; d < devices; ++d) {
      Slot slot; slot.gpu = d; slot.slab = s;
      slot.tiles.reserve((size_t)m_slabTiles);
      m_slots.push_back(slot);
    }
  // Tiles the caller may have queued ahead of the feeder: two slabs, whatever the number of GPUs.  Together with the
  // slabs in flight that bounds the caller-allocated tiles alive inside the queue to (2 + slots) * slabTiles plus the
  // finished ones the caller has not drained yet (INTEGRATION.md "Memory").
  m_inLimit = (size_t)m_slabTiles * 2;
  m_feeder = std::thread(&TileBatchQueue::feederMain, this);
  m_harvester = std::thread(&TileBatchQueue::harvesterMain, this);
  return true;
}

// ---- hand-over to the caller ------------------------------------------------------------------------
void TileBatchQueue::publish(std::vector<TileDataPtr>& tiles) noexcept
{
  {
    std::lock_guard<std::mutex> lock(m_doneMutex);
    for (TileDataPtr& t : tiles) m_done.push_back(t);
    m_pending -= (long long)tiles.size();
  }
  m_doneCv.notify_all();
}

void TileBatchQueue::drainDone() noexcept
{
  std::lock_guard<std::mutex> lock(m_doneMutex);
  while (!m_done.empty()) {
    getResultQueue().push(m_done.front());
    m_done.pop_front();
  }
}

// ---- caller's thread ----------------------------------------------------------------------------------
void TileBatchQueue::addTileData(TileDataPtr tileData) noexcept
{
  if (tileData == nullptr) return;
  if ((++m_addsSincePoll & 255u) == 0u) drainDone();
  const Encoding enc = tileData->getOptions().getEncoding();
  const bool gpuTile = tileData->isEncoding() && (enc == Encoding::BC1 || enc == Encoding::BC2 || enc == Encoding::BC3);
  if (!gpuTile) { runOnHost(tileData); return; }

  const int w = tileData->getWidth(), h = tileData->getHeight();
  const bool valid = tileData->getPaletteData() != nullptr && tileData->getIndexedData() != nullptr &&
                     tileData->getDeflatedData() != nullptr && tileData->getIndex() >= 0 && w > 0 && h > 0 && w <= 64 && h <= 64;
  if (valid && !m_gpus.empty() &&
      (pixelTypeOf(tileData->getOptions()) != m_type || tileData->getOptions().getEncodingQuality() != m_quality ||
       tileData->getOptions().isDeflate() != m_deflate)) {
    // the slabs are bound to the first tile's (type, level, framing); a tile that asks for something else
    // runs its own functor, which reaches the GPU one tile at a time through the converter plugin
    runOnHost(tileData);
    return;
  }
  if (!valid || !ensureContexts(tileData->getOptions())) {
    // the reference reports a failed tile through size 0 (graphics.cpp:1140 refuses to write it)
    tileData->setSize(0);
    getResultQueue().push(tileData);
    return;
  }
  { std::lock_guard<std::mutex> lock(m_doneMutex); ++m_pending; }
  {
    // "blocks execution as long as the input queue is full" (tilethreadpool_base.h:64)
    std::unique_lock<std::mutex> lock(m_inMutex);
    m_inSpaceCv.wait(lock, [this] { return m_inQueue.size() < m_inLimit || m_stop; });
    m_inQueue.push_back(tileData);
  }
  m_inCv.notify_one();
}

TileDataPtr TileBatchQueue::getResult() noexcept
{
  // slabs are harvested in submission order, so anything not yet published has a larger index than
  // the heap's top: the done list only needs a look when the heap is empty
  if (getResultQueue().empty()) drainDone();
  if (getResultQueue().empty()) return TileDataPtr();
  TileDataPtr top = getResultQueue().top();
  getResultQueue().pop();
  return top;
}

const TileDataPtr TileBatchQueue::peekResult() noexcept
{
  if (getResultQueue().empty()) drainDone();
  return getResultQueue().empty() ? TileDataPtr() : getResultQueue().top();
}

void TileBatchQueue::waitForResult() noexcept
{
  // the caller has nothing more to add right now: ship the partly filled slab, then block until
  // something is ready or the queue is idle
  { std::lock_guard<std::mutex> lock(m_inMutex); m_flush = true; }
  m_inCv.notify_one();
  drainDone();
  if (!getResultQueue().empty()) return;
  std::unique_lock<std::mutex> lock(m_doneMutex);
  m_doneCv.wait(lock, [this] { return !m_done.empty() || m_pending == 0; });
  while (!m_done.empty()) { getResultQueue().push(m_done.front()); m_done.pop_front(); }
}

bool TileBatchQueue::finished() noexcept
{
  drainDone();
  std::lock_guard<std::mutex> lock(m_doneMutex);
  return m_pending == 0 && m_done.empty() && getResultQueue().empty();
}

// ---- feeder: tiles -> pinned slabs -> GPU --------------------------------------------------------------
void TileBatchQueue::submitSlot(size_t slotIndex) noexcept
{
  Slot& slot = m_slots[slotIndex];
  if (slot.tiles.empty()) return;
  tcb200_ctx* ctx = m_gpus[(size_t)slot.gpu].ctx;
  { std::lock_guard<std::mutex> lock(m_flightMutex); slot.busy = true; }
  if (tcb200_slab_submit(ctx, slot.slab, (int)slot.tiles.size()) != TCB200_OK) {
    setError(tcb200_last_error(ctx));
    finishSlot(slot, false);
    { std::lock_guard<std::mutex> lock(m_flightMutex); slot.busy = false; }
    return;
  }
  { std::lock_guard<std::mutex> lock(m_flightMutex); m_inflight.push_back(slotIndex); }
  m_flightCv.notify_one();
}

void TileBatchQueue::feederMain() noexcept
{
  std::atomic<bool>& stop = m_stop;
  size_t current = 0;
  bool acquired = false;            // the feeder owns slot `current` (it is not in flight)
  std::deque<TileDataPtr> batch;
  for (;;) {
    bool flush = false;
    {
      std::unique_lock<std::mutex> lock(m_inMutex);
      m_inCv.wait(lock, [this] { return m_stop || m_flush || !m_inQueue.empty(); });
      if (m_stop) return;
      // take what fits the slab being filled, not the whole queue: the rest stays under the m_inLimit bound
      const size_t room = (size_t)m_slabTiles - (acquired ? m_slots[current].tiles.size() : 0);
      while (!m_inQueue.empty() && batch.size() < room) { batch.push_back(m_inQueue.front()); m_inQueue.pop_front(); }
      flush = m_flush && m_inQueue.empty();
      if (flush) m_flush = false;
    }
    m_inSpaceCv.notify_all();
    for (TileDataPtr& tile : batch) {
      Slot& slot = m_slots[current];
      if (!acquired) {
        // the slot may still be in flight from the previous lap (the analogue of "queue full");
        // the harvester empties slot.tiles before it clears `busy`
        std::unique_lock<std::mutex> lock(m_flightMutex);
        m_slotFreeCv.wait(lock, [&] { return !slot.busy || stop.load(); });
        if (stop.load()) return;
        acquired = true;
      }
      tcb200_ctx* ctx = m_gpus[(size_t)slot.gpu].ctx;
      uint8_t* in = tcb200_slab_input(ctx, slot.slab);
      uint16_t* sizes = tcb200_slab_sizes(ctx, slot.slab);
      if (in == nullptr || sizes == nullptr) {
        setError(tcb200_last_error(ctx));
        tile->setSize(0);
        std::vector<TileDataPtr> one(1, tile);
        publish(one);
        continue;
      }
      const int w = tile->getWidth(), h = tile->getHeight();
      const size_t i = slot.tiles.size();
      uint8_t* record = in + i * TCB200_TILE_RECORD;
      std::memcpy(record, tile->getPaletteData().get(), PALETTE_SIZE);
      std::memcpy(record + PALETTE_SIZE, tile->getIndexedData().get(), (size_t)w * h);
      sizes[2 * i] = (uint16_t)w; sizes[2 * i + 1] = (uint16_t)h;
      slot.tiles.push_back(tile);
      if ((int)slot.tiles.size() == m_slabTiles) { submitSlot(current); current = (current + 1) % m_slots.size(); acquired = false; }
    }
    batch.clear();
    if (flush && acquired && !m_slots[current].tiles.empty()) { submitSlot(current); current = (current + 1) % m_slots.size(); acquired = false; }
  }
}

// ---- harvester: GPU -> the tiles' own buffers ----------------------------------------------------------
void TileBatchQueue::finishSlot(Slot& slot, bool ok) noexcept
{
  tcb200_ctx* ctx = m_gpus[(size_t)slot.gpu].ctx;
  const uint8_t* out = ok ? tcb200_slab_output(ctx, slot.slab) : nullptr;
  const int stride = tcb200_record_stride(ctx);
  const size_t n = slot.tiles.size();
  auto finish = [&](size_t lo, size_t hi) {
    std::unique_ptr<RecordDeflater> deflater(m_deflate ? new RecordDeflater() : nullptr);
    for (size_t i = lo; i < hi; ++i) {
      TileData& tile = *slot.tiles[i];
      const int size = tcb200_record_size(m_type, tile.getWidth(), tile.getHeight());
      if (out == nullptr || size <= 0) { tile.setSize(0); continue; }
      const uint8_t* record = out + i * (size_t)stride;
      if (m_deflate) tile.setSize((int)deflater->run(record, (uint32_t)size, tile.getDeflatedData().get(), (uint32_t)size * 2));
      else { std::memcpy(tile.getDeflatedData().get(), record, (size_t)size); tile.setSize(size); }
    }
  };
  const unsigned workers = m_deflate ? std::min<unsigned>(m_threads, (unsigned)std::max<size_t>(1, n / 16)) : 1;
  if (workers <= 1) {
    finish(0, n);
  } else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < workers; ++t) pool.emplace_back(finish, n * t / workers, n * (t + 1) / workers);
    for (std::thread& th : pool) th.join();
  }
  publish(slot.tiles);
  slot.tiles.clear();
}

void TileBatchQueue::harvesterMain() noexcept
{
  std::atomic<bool>& stop = m_stop;
  for (;;) {
    size_t index;
    {
      std::unique_lock<std::mutex> lock(m_flightMutex);
      m_flightCv.wait(lock, [&] { return !m_inflight.empty() || stop.load(); });
      if (m_inflight.empty()) return;             // stop requested and nothing in flight
      index = m_inflight.front();
      m_inflight.pop_front();
    }
    Slot& slot = m_slots[index];
    tcb200_ctx* ctx = m_gpus[(size_t)slot.gpu].ctx;
    const bool ok = tcb200_slab_wait(ctx, slot.slab) == TCB200_OK;
    if (!ok) setError(tcb200_last_error(ctx));
    finishSlot(slot, ok);
    { std::lock_guard<std::mutex> lock(m_flightMutex); slot.busy = false; }
    m_slotFreeCv.notify_all();
  }
}

// ---- host lane: tiles the GPU path does not cover run their own functor on worker threads -----------
void TileBatchQueue::hostWorker() noexcept
{
  std::atomic<bool>& stop = m_stop;
  for (;;) {
    TileDataPtr tile;
    {
      std::unique_lock<std::mutex> lock(m_hostMutex);
      m_hostWake.wait(lock, [&] { return stop.load() || !m_hostQueue.empty(); });
      if (m_hostQueue.empty()) return;
      tile = m_hostQueue.front();
      m_hostQueue.pop_front();
      ++m_hostActive;
    }
    m_hostSpaceCv.notify_one();
    (*tile)();                                   // TileData::operator(): encode() or decode() on the CPU
    { std::lock_guard<std::mutex> lock(m_hostMutex); --m_hostActive; }
    std::vector<TileDataPtr> one(1, tile);
    publish(one);
  }
}

void TileBatchQueue::runOnHost(TileDataPtr tile) noexcept
{
  if (m_workers.empty())
    for (unsigned t = 0; t < m_threads; ++t) m_workers.emplace_back(&TileBatchQueue::hostWorker, this);
  { std::lock_guard<std::mutex> lock(m_doneMutex); ++m_pending; }
  {
    // bounded like the reference's input queue (getMaxTiles(), tilethreadpool_posix.cpp:66)
    std::unique_lock<std::mutex> lock(m_hostMutex);
    m_hostSpaceCv.wait(lock, [this] { return m_hostQueue.size() < (size_t)std::max(1u, getMaxTiles()); });
    m_hostQueue.push_back(tile);
  }
  m_hostWake.notify_one();
}

// ---- the two link-time factory functions of tilethreadpool_base.h:38,44 --------------------------
unsigned getThreadPoolAutoThreads()
{
  return std::max(1u, std::min(256u, std::thread::hardware_concurrency()));
}

ThreadPoolPtr createThreadPool(int threadNum, int tileNum)
{
  return ThreadPoolPtr(new TileBatchQueue((unsigned)std::max(1, threadNum), (unsigned)std::max(1, tileNum)));
}

}  // namespace tc
