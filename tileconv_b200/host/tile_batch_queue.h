// tileconv_b200/host/tile_batch_queue.h -- the pinned-memory batch queue that takes the place of
// TileThreadPoolPosix / TileThreadPoolWin32 (reference tileconv/tilethreadpool_posix.cpp:39-159,
// tilethreadpool_win32.cpp) behind the unchanged TileThreadPool interface
// (tilethreadpool_base.h:47-115) and the unchanged factory function createThreadPool()
// (tilethreadpool_base.h:44; chosen at link time, tilethreadpool.h:25-29).
//
// Instead of N workers each running one TileData functor at a time with 50 ms polling sleeps
// (tilethreadpool_posix.cpp:66-68,101-103,140-143), the caller's thread only hands tiles over; a
// FEEDER thread copies them into pinned slabs and submits every full slab (upload, kernel and
// download run on that slab's own CUDA stream while the next slab fills: double buffering, three
// slabs per GPU), and a HARVESTER thread waits for slabs in submission order, copies the records
// back into the tiles' own buffers and publishes them.  With several GPUs consecutive slabs go to
// consecutive devices, i.e. the tile-index range is sharded slab by slab and no data ever moves
// between GPUs.  Results reach the caller through the base class's index-ordered result heap
// (touched by the caller's thread only, because hasResult() reads it without a lock), so callers
// that drain with peekResult()->getIndex() == next (graphics.cpp:125-143) work unchanged.
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

#include <atomic>
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
  // threadNum: host threads for the zlib step and the host lane; tileNum: the caller's window
  // (graphics.cpp:52 passes 64) -- kept as getMaxTiles() but NOT used as the batch size, which is
  // slabTiles (TCB200_SLAB_TILES).  maxGpus 0 = all visible devices (TCB200_GPUS).
  TileBatchQueue(unsigned threadNum, unsigned tileNum, int slabTiles = 4096, int maxGpus = 0) noexcept;
  ~TileBatchQueue() noexcept;

  void addTileData(TileDataPtr tileData) noexcept;
  TileDataPtr getResult() noexcept;
  const TileDataPtr peekResult() noexcept;
  void waitForResult() noexcept;
  bool finished() noexcept;

  int gpuCount() const noexcept { return (int)m_gpus.size(); }
  std::string lastError() noexcept;

protected:
  void threadActivated() noexcept {}
  void threadDeactivated() noexcept {}
  int getActiveThreads() noexcept;

private:
  struct Gpu { tcb200_ctx* ctx = nullptr; };
  struct Slot { int gpu = 0, slab = 0; bool busy = false; std::vector<TileDataPtr> tiles; };

  bool ensureContexts(const Options& options) noexcept;
  void setError(const std::string& msg) noexcept;
  void publish(std::vector<TileDataPtr>& tiles) noexcept;      // any thread: hand finished tiles to the caller
  void drainDone() noexcept;                                    // caller's thread: done list -> result heap
  void feederMain() noexcept;
  void harvesterMain() noexcept;
  void submitSlot(size_t slotIndex) noexcept;                   // feeder
  void finishSlot(Slot& slot, bool ok) noexcept;                // harvester: copy / deflate records into the tiles
  void hostWorker() noexcept;
  void runOnHost(TileDataPtr tile) noexcept;

  unsigned m_threads;
  int m_slabTiles, m_maxGpus;
  int m_type = 0, m_quality = -1;
  bool m_deflate = false;
  std::vector<Gpu> m_gpus;
  std::vector<Slot> m_slots;         // gpus x slabs, used round-robin by the feeder

  // caller -> feeder
  std::mutex m_inMutex;
  std::condition_variable m_inCv, m_inSpaceCv;
  std::deque<TileDataPtr> m_inQueue;
  bool m_flush = false;
  std::atomic<bool> m_stop{false};   // set once, by the destructor
  size_t m_inLimit = 0;
  // feeder -> harvester, and slot recycling
  std::mutex m_flightMutex;
  std::condition_variable m_flightCv, m_slotFreeCv;
  std::deque<size_t> m_inflight;     // submitted slots, oldest first
  // anyone -> caller
  std::mutex m_doneMutex;
  std::condition_variable m_doneCv;
  std::deque<TileDataPtr> m_done;
  long long m_pending = 0;           // tiles accepted and not yet in m_done (guarded by m_doneMutex)
  std::string m_error;               // guarded by m_doneMutex
  // host lane
  std::mutex m_hostMutex;
  std::condition_variable m_hostWake, m_hostSpaceCv;
  std::deque<TileDataPtr> m_hostQueue;
  int m_hostActive = 0;

  std::thread m_feeder, m_harvester;
  std::vector<std::thread> m_workers;
  unsigned m_addsSincePoll = 0;
};

}  // namespace tc

#endif
