#include "tile_batch_queue.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

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

// one-shot zlib stream at level 9 into a buffer of 2x the source size (compress.cpp:36,54-67;
// tiledata.cpp:142-143); 0 on failure
uint32_t deflateRecord(const uint8_t* src, uint32_t srcSize, uint8_t* dst, uint32_t dstSize)
{
  z_stream z;
  std::memset(&z, 0, sizeof(z));
  if (deflateInit(&z, 9) != Z_OK) return 0;
  z.avail_in = srcSize; z.next_in = const_cast<Bytef*>(src);
  z.avail_out = dstSize; z.next_out = dst;
  const int rc = ::deflate(&z, Z_FINISH);
  const uint32_t written = dstSize - z.avail_out;
  deflateEnd(&z);
  return rc == Z_STREAM_ERROR ? 0 : written;
}

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
  while (!m_inflight.empty()) harvestOldest();
  for (Gpu& g : m_gpus) tcb200_destroy(g.ctx);
  {
    std::lock_guard<std::mutex> lock(m_hostMutex);
    m_hostStop = true;
  }
  m_hostWake.notify_all();
  for (std::thread& t : m_workers) t.join();
}

// ---- host lane: tiles the GPU path does not cover run their own functor on worker threads ----------
void TileBatchQueue::hostWorker() noexcept
{
  for (;;) {
    TileDataPtr tile;
    {
      std::unique_lock<std::mutex> lock(m_hostMutex);
      m_hostWake.wait(lock, [this] { return m_hostStop || !m_hostQueue.empty(); });
      if (m_hostQueue.empty()) return;          // stop requested and nothing left
      tile = m_hostQueue.front();
      m_hostQueue.pop_front();
      ++m_hostActive;
    }
    (*tile)();                                   // TileData::operator(): encode() or decode() on the CPU
    {
      std::lock_guard<std::mutex> lock(m_hostMutex);
      m_hostDone.push_back(tile);
      --m_hostActive;
    }
    m_hostDoneCv.notify_all();
  }
}

void TileBatchQueue::runOnHost(TileDataPtr tile) noexcept
{
  if (m_workers.empty())
    for (unsigned t = 0; t < m_threads; ++t) m_workers.emplace_back(&TileBatchQueue::hostWorker, this);
  {
    // bounded like the reference's input queue (getMaxTiles(), tilethreadpool_posix.cpp:66)
    std::unique_lock<std::mutex> lock(m_hostMutex);
    m_hostDoneCv.wait(lock, [this] { return m_hostQueue.size() < (size_t)std::max(1u, getMaxTiles()); });
    m_hostQueue.push_back(tile);
  }
  m_hostWake.notify_one();
}

void TileBatchQueue::drainHostResults() noexcept
{
  std::lock_guard<std::mutex> lock(m_hostMutex);
  while (!m_hostDone.empty()) {
    getResultQueue().push(m_hostDone.front());
    m_hostDone.pop_front();
  }
}

bool TileBatchQueue::ensureContexts(const Options& options) noexcept
{
  if (!m_gpus.empty()) return true;
  m_type = pixelTypeOf(options);
  m_quality = options.getEncodingQuality();
  m_deflate = options.isDeflate();
  int devices = tcb200_device_count();
  if (m_maxGpus > 0) devices = std::min(devices, m_maxGpus);
  if (devices <= 0 || m_type == 0) {
    m_error = devices <= 0 ? "no CUDA device (there is no CPU fallback)" : "pixel type is not BC1/BC2/BC3";
    return false;
  }
  for (int d = 0; d < devices; ++d) {
    tcb200_ctx* ctx = tcb200_create(d, m_type, m_quality, m_slabTiles);
    if (ctx == nullptr) {
      m_error = tcb200_last_error(nullptr);
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
      m_slots.push_back(slot);
    }
  m_current = 0;
  return true;
}

void TileBatchQueue::failTile(TileDataPtr tile) noexcept
{
  // the reference reports a failed tile through size 0 (graphics.cpp:1140 refuses to write it)
  tile->setSize(0);
  getResultQueue().push(tile);
}

void TileBatchQueue::addTileData(TileDataPtr tileData) noexcept
{
  if (tileData == nullptr) return;
  poll();
  {
    const Encoding enc = tileData->getOptions().getEncoding();
    const bool gpuTile = tileData->isEncoding() && (enc == Encoding::BC1 || enc == Encoding::BC2 || enc == Encoding::BC3);
    if (!gpuTile) { runOnHost(tileData); return; }
  }
  const int w = tileData->getWidth(), h = tileData->getHeight();
  const bool valid = tileData->isEncoding() && tileData->getPaletteData() != nullptr && tileData->getIndexedData() != nullptr &&
                     tileData->getDeflatedData() != nullptr && tileData->getIndex() >= 0 && w > 0 && h > 0 && w <= 64 && h <= 64;
  if (!valid || !ensureContexts(tileData->getOptions())) { failTile(tileData); return; }

  Slot* slot = &m_slots[m_current];
  // the slot being filled may still be in flight from the previous lap: wait for it (this is the
  // only place the producer blocks, the analogue of "queue full" in tilethreadpool_posix.cpp:66)
  while (std::find(m_inflight.begin(), m_inflight.end(), m_current) != m_inflight.end()) harvestOldest();

  tcb200_ctx* ctx = m_gpus[(size_t)slot->gpu].ctx;
  uint8_t* in = tcb200_slab_input(ctx, slot->slab);
  uint16_t* sizes = tcb200_slab_sizes(ctx, slot->slab);
  if (in == nullptr || sizes == nullptr) { m_error = tcb200_last_error(ctx); failTile(tileData); return; }
  const size_t i = slot->tiles.size();
  uint8_t* record = in + i * TCB200_TILE_RECORD;
  std::memcpy(record, tileData->getPaletteData().get(), PALETTE_SIZE);
  std::memcpy(record + PALETTE_SIZE, tileData->getIndexedData().get(), (size_t)w * h);
  sizes[2 * i] = (uint16_t)w; sizes[2 * i + 1] = (uint16_t)h;
  slot->tiles.push_back(tileData);
  if ((int)slot->tiles.size() == m_slabTiles) submitCurrent();
}

void TileBatchQueue::submitCurrent() noexcept
{
  if (m_slots.empty()) return;
  Slot& slot = m_slots[m_current];
  if (slot.tiles.empty()) return;
  tcb200_ctx* ctx = m_gpus[(size_t)slot.gpu].ctx;
  if (tcb200_slab_submit(ctx, slot.slab, (int)slot.tiles.size()) != TCB200_OK) {
    m_error = tcb200_last_error(ctx);
    harvest(slot, false);
  } else {
    m_inflight.push_back(m_current);
  }
  m_current = (m_current + 1) % m_slots.size();
}

void TileBatchQueue::harvest(Slot& slot, bool ok) noexcept
{
  tcb200_ctx* ctx = m_gpus[(size_t)slot.gpu].ctx;
  const uint8_t* out = ok ? tcb200_slab_output(ctx, slot.slab) : nullptr;
  const int stride = tcb200_record_stride(ctx);
  const size_t n = slot.tiles.size();
  auto finish = [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) {
      TileData& tile = *slot.tiles[i];
      const int size = tcb200_record_size(m_type, tile.getWidth(), tile.getHeight());
      if (out == nullptr || size <= 0) { tile.setSize(0); continue; }
      const uint8_t* record = out + i * (size_t)stride;
      if (m_deflate) tile.setSize((int)deflateRecord(record, (uint32_t)size, tile.getDeflatedData().get(), (uint32_t)size * 2));
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
  for (TileDataPtr& tile : slot.tiles) getResultQueue().push(tile);
  slot.tiles.clear();
}

void TileBatchQueue::harvestOldest() noexcept
{
  if (m_inflight.empty()) return;
  Slot& slot = m_slots[m_inflight.front()];
  m_inflight.pop_front();
  tcb200_ctx* ctx = m_gpus[(size_t)slot.gpu].ctx;
  const bool ok = tcb200_slab_wait(ctx, slot.slab) == TCB200_OK;
  if (!ok) m_error = tcb200_last_error(ctx);
  harvest(slot, ok);
}

void TileBatchQueue::poll() noexcept
{
  drainHostResults();
  while (!m_inflight.empty()) {
    Slot& slot = m_slots[m_inflight.front()];
    if (tcb200_slab_ready(m_gpus[(size_t)slot.gpu].ctx, slot.slab) != 1) break;
    harvestOldest();
  }
}

TileDataPtr TileBatchQueue::getResult() noexcept
{
  poll();
  if (getResultQueue().empty()) return TileDataPtr();
  TileDataPtr top = getResultQueue().top();
  getResultQueue().pop();
  return top;
}

const TileDataPtr TileBatchQueue::peekResult() noexcept
{
  poll();
  return getResultQueue().empty() ? TileDataPtr() : getResultQueue().top();
}

void TileBatchQueue::waitForResult() noexcept
{
  // the caller has nothing more to add right now: ship the partly filled slab, then block until
  // something is ready or the queue is idle
  submitCurrent();
  drainHostResults();
  if (!getResultQueue().empty()) return;
  if (!m_inflight.empty()) { harvestOldest(); return; }
  std::unique_lock<std::mutex> lock(m_hostMutex);
  m_hostDoneCv.wait(lock, [this] { return !m_hostDone.empty() || (m_hostQueue.empty() && m_hostActive == 0); });
  while (!m_hostDone.empty()) { getResultQueue().push(m_hostDone.front()); m_hostDone.pop_front(); }
}

bool TileBatchQueue::finished() noexcept
{
  poll();
  const bool filling = !m_slots.empty() && !m_slots[m_current].tiles.empty();
  bool hostIdle;
  {
    std::lock_guard<std::mutex> lock(m_hostMutex);
    hostIdle = m_hostQueue.empty() && m_hostActive == 0 && m_hostDone.empty();
  }
  return !filling && m_inflight.empty() && hostIdle && getResultQueue().empty();
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
