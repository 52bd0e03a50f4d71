// tileconv_b200/host/slab_pipeline.h -- the streaming fast path of the stand-alone driver (tbc_encode).
//
// The reference's drivers (Graphics::tisToTBC / mosToMBC, tileconv/graphics.cpp:66-169, 278-392) build one TileData
// with three heap buffers per tile and hand it to the pool; at GPU rates that caller loop, not the encoder, is the
// limit.  A TIS body is already an array of 5120-byte tile records, so this path skips the per-tile
// objects: worker threads pread() tile records straight into the contexts' pinned input slabs (tcb200_slab_input),
// slabs are submitted round-robin over the devices, and finished records go from the pinned output slabs into the
// output file with the reference's framing (u32 size + payload per tile, tileconv/graphics.cpp:1138-1165), deflated
// first at level 9 when -u is not given (tileconv/compress.cpp:36,54-67).  Same bytes as the TileData route; the
// tests compare both with each other and with the reference CLI.
#ifndef TCB200_SLAB_PIPELINE_H
#define TCB200_SLAB_PIPELINE_H

#include <cstdint>
#include <string>

#include "tc_iface.h"

namespace tc {

// Where the tiles come from.  fill() may be called from several threads at once on disjoint ranges.
class TileFeed
{
public:
  virtual ~TileFeed() {}
  virtual unsigned count() const noexcept = 0;
  // tiles [first, first + n) -> n records of 5120 bytes (1024 B palette, then w*h indices at pitch w) and n x {w,h}
  virtual bool fill(unsigned first, unsigned n, uint8_t* records, uint16_t* wh) noexcept = 0;
  // size of tile i (known from the container's geometry alone); the default is a full 64x64 tile
  virtual void dims(unsigned i, uint16_t& w, uint16_t& h) const noexcept { (void)i; w = 64; h = 64; }
};

// Seconds of wall time each stage was busy on the thread that drives it (stages overlap, so they do not add up to
// `total`), for the steady-state report of tools/cli_rate.py.
struct PipelineTimes
{
  double setup = 0;      // contexts, pinned slabs, worker threads (CUDA start-up lives here)
  double fill = 0;       // feeder: input file / MOS image -> pinned input slabs (parallel pread / gather)
  double submit = 0;     // feeder: tcb200_slab_submit (enqueue H2D + kernel + D2H)
  double gpuWait = 0;    // harvester: blocked in tcb200_slab_wait
  double frame = 0;      // harvester: size prefixes / deflate into the staging buffer (parallel)
  double write = 0;      // harvester: pwrite of the framed records (parallel)
  double total = 0;      // first fill to last write
  unsigned long long tiles = 0, bytesIn = 0, bytesOut = 0;
  int devices = 0, slabTiles = 0, threads = 0;
};

// Encodes every tile of `feed` with options' -t / -q / -u and appends the framed tiles to file descriptor `fd` at
// byte offset `offset` (the container header is the caller's).  false + message on any failure.
bool runSlabPipeline(const Options& options, TileFeed& feed, int fd, unsigned long long offset, std::string& error,
                     PipelineTimes* times) noexcept;

// True when the fast path handles this job (a BC1/2/3 encode and not disabled with TCB200_TILEDATA_ROUTE=1).
bool slabPipelineApplies(const Options& options) noexcept;

}  // namespace tc

#endif
