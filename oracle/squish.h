/* oracle/squish.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Stand-in for the <squish.h> of libsquish, the third-party library the reference links as
 * -lsquish (reference tileconv/Makefile:10, tileconv/config.mk:6,12) and which is NOT vendored in
 * /root/reference (tileconv/squish/remove.me is the only file there).  It declares exactly the
 * names the reference's call sites use:
 *   #include <squish.h>                       tileconv/converter_dxt.cpp:23
 *   squish::kDxt1 / kDxt3 / kDxt5             tileconv/converter_dxt.cpp:185-187
 *   squish::kColourRangeFit / kColourClusterFit / kWeightColourByAlpha /
 *   squish::kColourIterativeClusterFit        tileconv/converter_dxt.cpp:195-204
 *   squish::Compress(buffer, dst, flags)      tileconv/converter_dxt.cpp:221
 * so that the reference's own translation units compile and link against the restated
 * implementation in squish_oracle.cpp unchanged (recipe: oracle/Makefile, target _ref).
 *
 * PARITY UNPINNED: libsquish is un-pinned by the reference (no version, no lockfile, no tests, no
 * golden vectors -- SURVEY.md F1/F2), so the arithmetic here restates libsquish 1.15's published
 * algorithm with the fixed choices listed in squish_oracle.cpp.
 */
#ifndef TCB200_ORACLE_SQUISH_H
#define TCB200_ORACLE_SQUISH_H

namespace squish {

typedef unsigned char u8;

enum
{
  kDxt1                      = (1 << 0),
  kDxt3                      = (1 << 1),
  kDxt5                      = (1 << 2),
  kColourClusterFit          = (1 << 3),
  kColourRangeFit            = (1 << 4),
  kWeightColourByAlpha       = (1 << 7),
  kColourIterativeClusterFit = (1 << 8)
};

/* rgba: 16 pixels, memory order R,G,B,A, row-major 4x4.  block: 8 bytes (kDxt1) or 16 bytes. */
void Compress(u8 const* rgba, void* block, int flags, float* metric = 0);
void CompressMasked(u8 const* rgba, int mask, void* block, int flags, float* metric = 0);

}  // namespace squish

/* Instrumentation (oracle only): per-thread counters of the work the fitters actually did.  They
 * feed the flop model of the ALU roofline (DESIGN.md "Algorithmic work").  */
struct SquishOracleCounters
{
  unsigned long long blocks;          /* Compress calls */
  unsigned long long single_blocks;   /* SingleColourFit */
  unsigned long long range_blocks;    /* RangeFit (incl. the empty-set case) */
  unsigned long long cluster_blocks;  /* ClusterFit */
  unsigned long long cand3;           /* (i,j) candidates evaluated by 3-colour cluster sweeps */
  unsigned long long cand4;           /* (i,j,k) candidates evaluated by 4-colour cluster sweeps */
  unsigned long long sweeps3;         /* orderings swept, 3-colour */
  unsigned long long sweeps4;         /* orderings swept, 4-colour */
  unsigned long long points;          /* sum of ColourSet counts over cluster blocks */
};
extern "C" void squish_oracle_counters_reset(void);
extern "C" void squish_oracle_counters_get(SquishOracleCounters* out);

#endif
