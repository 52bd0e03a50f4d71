// tileconv_b200/host/tile_batch_queue.h -- the pinned-memory batch queue that takes the place of
// TileThreadPoolPosix / TileThreadPoolWin32 (reference tileconv/tilethreadpool_posix.cpp:39-159,
// tilethreadpool_win32.cpp) behind the unchanged TileThreadPool interface
// (tilethreadpool_base.h:47-115) and the unchanged factory function createThreadPool()
// (tilethreadpool_base.h:44; chosen at link time, tilethreadpool.h:25-29).
//
// Instead of N workers each running one TileData functor at a time with 50 ms polling sleeps
// (tilethreadpool_posix.cpp:66-68,101-103,140-143), tiles are copied into pinned slabs; a full slab
// is uploaded, encoded and downloaded on its own CUDA stream while the next slab fills (double
// buffering; three slabs per GPU).  With several GPUs consecutive slabs go to consecutive devices,
// i.e. the tile-index range is sharded slab by slab and no data ever moves between GPUs.  Results
// come back through the base class's index-ordered result heap, so callers that drain with
// peekResult()->getIndex() == next (graphics.cpp:125-143) work unchanged.
//
// Tiles that are not a BC1/BC2/BC3 ENCODE -- the decode direction (tbcToTIS / mbcToMOS feed the same
// pool, graphics.cpp:172-275,395-517) and the raw passthrough type 0 -- are outside the GPU path: they
// run the tile's own functor (TileData::operator(), tiledata.cpp:58-66) on `threadNum` host worker
// threads, exactly what the reference's pool does with every tile, so the queue stays a complete
// replacement for createThreadPool()'s object.
//
// zlib framing (TileData::encode's deflate step, tiledata.cpp:139-148) stays on the host: when
// Options::isDeflate() is set the harvested records are deflated (level 9, compress.cpp:36) by
// `threadNum` host threads before they are published.
#ifndef TCB200_TILE_BATCH_QUEUE_H
#define TCB200_TILE_BATCH_QUEUE_H

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "tc_iface.h"

struct tcb200_ctx;

namespace tc {

class TileBatchQueue : public TileThreadPool
{
public:
  // threadNum: host threads for the zlib step; tileNum: the caller's window (graphics.cpp:52 passes
  // 64) -- kept as getMaxTiles() but NOT used as the batch size, which is slabTiles.
  TileBatchQueue(unsigned threadNum, unsigned tileNum, int slabTiles = 4096, int maxGpus = 0) noexcept;
  ~TileBatchQueue() noexcept;

  void addTileData(TileDataPtr tileData) noexcept;
  TileDataPtr getResult() noexcept;
  const TileDataPtr peekResult() noexcept;
  void waitForResult() noexcept;
  bool finished() noexcept;

  int gpuCount() const noexcept { return (int)m_gpus.size(); }
  const std::string& lastError() const noexcept { return m_error; }

protected:
  void threadActivated() noexcept {}
  void threadDeactivated() noexcept {}
  int getActiveThreads() noexcept { return (int)m_inflight.size(); }

private:
  struct Gpu { tcb200_ctx* ctx = nullptr; };
  struct Slot { int gpu = 0, slab = 0; std::vector<TileDataPtr> tiles; };

  bool ensureContexts(const Options& options) noexcept;
  void submitCurrent() noexcept;
  void harvest(Slot& slot, bool ok) noexcept;
  void harvestOldest() noexcept;     // blocking
  void poll() noexcept;              // non-blocking: publish every slab that has completed, in order
  void failTile(TileDataPtr tile) noexcept;
  void runOnHost(TileDataPtr tile) noexcept;   // non-GPU tiles: the reference's per-tile functor on a worker
  void hostWorker() noexcept;
  void drainHostResults() noexcept;            // main thread: move finished host tiles into the result heap

  unsigned m_threads;
  int m_slabTiles, m_maxGpus;
  int m_type = 0, m_quality = -1;
  bool m_deflate = false;
  std::vector<Gpu> m_gpus;
  std::vector<Slot> m_slots;         // gpus x slabs, used round-robin
  size_t m_current = 0;              // slot being filled
  std::deque<size_t> m_inflight;     // submitted slots, oldest first
  std::string m_error;
  // host side lane (decode / raw tiles)
  std::vector<std::thread> m_workers;
  std::mutex m_hostMutex;
  std::condition_variable m_hostWake, m_hostDoneCv;
  std::deque<TileDataPtr> m_hostQueue, m_hostDone;
  int m_hostActive = 0;
  bool m_hostStop = false;
};

}  // namespace tc

#endif
