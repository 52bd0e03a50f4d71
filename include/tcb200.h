/* include/tcb200.h -- C ABI of libtcb200.so: the B200 (sm_100a) implementation of tileconv's
 * TIS/MOS -> TBC/MBC encode hot path (palette expand + BC1/BC2/BC3 block fitting).
 *
 * Plain C: pointers and sizes only, no C++ or torch types.  Return convention follows the
 * reference's (everything there is noexcept and reports 0 / nullptr / false):
 * functions returning int give 0 on success and a negative TCB200_E_* code on failure, never throw;
 * the message is available from tcb200_last_error().  There is NO CPU fallback: without a usable
 * CUDA device every compute entry point fails with TCB200_E_NODEVICE.
 *
 * The reference has no FFI of its own (single C++ binary); each entry point below names the
 * reference interface it stands in for, so that a maintainer can bind it from the reference's
 * Converter / TileThreadPool seams.
 *
 * Threading: a context may be driven by one thread at a time, with one exception the batch queue relies
 * on: the slab functions of DIFFERENT slabs may run on different threads concurrently (one thread fills and
 * submits, another waits and reads), as long as each slab is handed from one thread to the other with a
 * proper happens-before edge.  Different contexts are independent.
 *
 * Data layouts (all little-endian, identical to the reference's in-memory tiles):
 *   input tile record  = 5120 bytes: 1024 B palette (256 x {B,G,R,x}) then 4096 B index area.
 *                        A w x h tile (w,h <= 64) keeps its indices in the first w*h bytes with
 *                        row pitch w  (tileconv/graphics.cpp:101-119 TIS, :322-342 MOS).
 *   output tile record = `stride` bytes, stride = tcb200_record_size(type,64,64): u16 w, u16 h,
 *                        then ceil(w/4)*ceil(h/4) blocks in row-major block order, 8 B (BC1) or
 *                        16 B (BC2/BC3) each  (tileconv/converter_dxt.cpp:109-144, FORMATS:36-41).
 *                        Only the first tcb200_record_size(type,w,h) bytes of a record are defined.
 */
#ifndef TCB200_H
#define TCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCB200_TILE_RECORD 5120

#define TCB200_OK            0
#define TCB200_E_ARG       (-1)   /* bad argument (null pointer, type/quality/size out of range) */
#define TCB200_E_NODEVICE  (-2)   /* no CUDA device / device index out of range */
#define TCB200_E_CUDA      (-3)   /* a CUDA call failed; see tcb200_last_error() */
#define TCB200_E_NOMEM     (-4)   /* allocation failed */

typedef struct tcb200_ctx tcb200_ctx;

/* Number of usable CUDA devices (0 when there is none).  Stands in for getThreadPoolAutoThreads()
 * (tileconv/tilethreadpool_base.h:38): "how wide can the fan-out be". */
int tcb200_device_count(void);

/* getRequiredSpace(w,h) + HEADER_TILE_ENCODED_SIZE for pixel type 1,2,3; 0 on error
 * (tileconv/converter_dxt.cpp:41-62, tileconv/tiledata.cpp:123). */
int tcb200_record_size(int type, int width, int height);

/* One encoder context = one device, one pixel type (-t 1/2/3 = BC1/BC2/BC3,
 * tileconv/options.cpp:95-110) and one encode quality (-q second digit 0..9, mapped onto
 * RangeFit / ClusterFit / weighted / iterative exactly as ConverterDxt::getFlags does,
 * tileconv/converter_dxt.cpp:179-212).  max_batch_tiles sizes the device-side double buffers
 * used by tcb200_encode (0 = default).  Replaces "new ConverterDxt per tile"
 * (tileconv/tiledata.cpp:114-121) + createThreadPool (tileconv/tilethreadpool_base.h:44).
 * Returns NULL on failure (message via tcb200_last_error(NULL)). */
tcb200_ctx* tcb200_create(int device, int type, int quality, int max_batch_tiles);
void        tcb200_destroy(tcb200_ctx* ctx);

/* libsquish-version knobs.  The reference links whatever libsquish the builder dropped into ./squish
 * (tileconv/Makefile:10, tileconv/squish/remove.me) and calls squish::Compress(rgba, block, flags)
 * without a metric (tileconv/converter_dxt.cpp:221), so the arithmetic behind that call depends on the
 * libsquish build a deployment used.  tcb200_create() fixes the choice "libsquish 1.15, scalar
 * simd_float.h": metric (1,1,1), Reciprocal = 1.0f/x.  tcb200_create_ex() can match the other builds:
 *   metric      weights of the colour error; (1,1,1) is what libsquish >= 1.11 uses for metric == NULL,
 *               (0.2126, 0.7152, 0.0722) is the kColourMetricPerceptual default of libsquish <= 1.10.
 *   reciprocal  TCB200_RECIP_IEEE: scalar build (SQUISH_USE_SSE 0).  TCB200_RECIP_SSE: simd_sse.h
 *               semantics (SQUISH_USE_SSE 1/2) -- Reciprocal = rcpps estimate + one Newton step in the
 *               power iteration and the cluster solve, Min/Max like minps/maxps, Truncate like
 *               cvttps2dq.  The estimate is measured from the HOST CPU's rcpps at create time, i.e. the
 *               result equals a libsquish SSE build running on this machine.
 * opt == NULL is tcb200_create().  Measured distance between the variants: profiles/r2_variant_sensitivity.md. */
#define TCB200_RECIP_IEEE 0
#define TCB200_RECIP_SSE  1
typedef struct tcb200_options
{
  uint32_t struct_size;      /* sizeof(tcb200_options) */
  float    metric[3];        /* r, g, b weights, finite and >= 0 */
  int32_t  reciprocal;       /* TCB200_RECIP_* */
} tcb200_options;
void        tcb200_default_options(tcb200_options* opt);
tcb200_ctx* tcb200_create_ex(int device, int type, int quality, int max_batch_tiles, const tcb200_options* opt);
int         tcb200_get_options(const tcb200_ctx* ctx, tcb200_options* opt);

/* Last error message of this context (or of the failed tcb200_create when ctx == NULL). */
const char* tcb200_last_error(const tcb200_ctx* ctx);

int tcb200_type(const tcb200_ctx* ctx);
int tcb200_quality(const tcb200_ctx* ctx);
int tcb200_device(const tcb200_ctx* ctx);
int tcb200_record_stride(const tcb200_ctx* ctx);   /* = tcb200_record_size(type, 64, 64) */

/* Host-to-host batch encode: n tile records in `tiles` (host memory, ideally pinned), optional
 * per-tile sizes `wh` (n x {u16 w, u16 h}; NULL = all 64x64), n output records written to `out`
 * (host).  Internally the batch is cut into chunks that are uploaded, encoded and downloaded on
 * separate CUDA streams with double buffering; returns when `out` is complete.
 * This is the batched form of Converter::convert(palette, indexed, encoded, w, h)
 * (tileconv/converter.h:85) as driven by TileData::encode (tileconv/tiledata.cpp:111-162, -u path). */
int tcb200_encode(tcb200_ctx* ctx, const uint8_t* tiles, const uint16_t* wh, int n, uint8_t* out);

/* Multi-GPU form of tcb200_encode, the in-process replacement of the reference's fan-out object
 * (createThreadPool + TileThreadPool, tileconv/tilethreadpool_base.h:38-81, tilethreadpool_posix.cpp:33-36):
 * ctxs[0..n_ctx) are contexts on DIFFERENT devices created with the same type, quality and options.  The tile
 * index range is cut into n_ctx contiguous shards, [g*n/n_ctx, (g+1)*n/n_ctx) goes to ctxs[g]; every device
 * uploads, encodes and downloads its shard on its own streams (one host thread per context inside the call)
 * and writes its records straight into out + first_tile*stride.  No collective, no reordering: records have a
 * fixed stride and tiles are independent.  Returns when `out` is complete; the first failing
 * shard's code, message via tcb200_last_error(ctxs[0]).  Pinned host buffers (tcb200_host_alloc) are needed
 * for the copies to overlap the kernels. */
int tcb200_encode_multi(tcb200_ctx* const* ctxs, int n_ctx, const uint8_t* tiles, const uint16_t* wh, int n, uint8_t* out);

/* Device-resident variant: all three pointers are device pointers on ctx's device; the kernel is
 * enqueued on `stream` (a cudaStream_t passed as void*, NULL = the context's compute stream) and
 * the call returns without synchronising.  d_tiles must be 16-byte aligned (the kernel stages records
 * with 128-bit loads), d_out 4-byte aligned. */
int tcb200_encode_device(tcb200_ctx* ctx, const void* d_tiles, const void* d_wh, int n, void* d_out, void* stream);

/* d_wh lives in device memory, so tcb200_encode_device cannot validate it up front the way the host entry
 * points do: the kernel rejects a record whose w or h is outside 1..64 (it writes a zero {w,h} header and
 * no blocks, the rest of that output record is untouched) and counts it.  This call waits for the
 * context's own streams and returns the number of records rejected since the context was created
 * (0 = none), negative on error.  Work enqueued on a caller-supplied stream must be synchronised by the caller first. */
long long tcb200_rejected_tiles(tcb200_ctx* ctx);

/* Asynchronous slab interface used by the batch queue (host side of TileThreadPool's role,
 * tileconv/tilethreadpool_base.h:64-81): the context owns `slabs` pinned input/output slabs of
 * max_batch_tiles records.  Fill tcb200_slab_input(), submit, later wait and read
 * tcb200_slab_output().  Slabs progress independently (upload, kernel and download of different
 * slabs overlap). */
int      tcb200_slab_count(const tcb200_ctx* ctx);
int      tcb200_slab_capacity(const tcb200_ctx* ctx);
uint8_t* tcb200_slab_input(tcb200_ctx* ctx, int slab);     /* capacity x 5120 B, pinned */
uint16_t* tcb200_slab_sizes(tcb200_ctx* ctx, int slab);    /* capacity x {w,h}, pinned  */
uint8_t* tcb200_slab_output(tcb200_ctx* ctx, int slab);    /* capacity x stride B, pinned */
int      tcb200_slab_submit(tcb200_ctx* ctx, int slab, int n);
int      tcb200_slab_wait(tcb200_ctx* ctx, int slab);      /* blocks until that slab's output is ready */
int      tcb200_slab_ready(tcb200_ctx* ctx, int slab);     /* 1 ready, 0 in flight, <0 error */

/* Pinned host memory for callers that want zero-copy-staging uploads (TileData's buffers are
 * plain new[] in the reference, tileconv/graphics.cpp:101-103). */
void* tcb200_host_alloc(size_t bytes);
void  tcb200_host_free(void* p);

/* Colors::palToARGB (tileconv/colors.cpp:42-59) for n 64x64 tile records: writes n x 4096 pixels
 * of 4 bytes in memory order B,G,R,A to host memory `bgra`.  Exposed for parity tests of the
 * expand stage; the encode kernels expand in shared memory and never write pixels to HBM. */
int tcb200_pal_to_argb(tcb200_ctx* ctx, const uint8_t* tiles, int n, uint8_t* bgra);

/* Decode direction, front half only ("next" row N3): n encoded tile records (stride =
 * tcb200_record_stride, each starting with its own u16 w, u16 h) -> n x 16384 bytes of pixels in
 * memory order B,G,R,A at row pitch w (the first w*h*4 bytes of each 16 KiB slot), i.e. what
 * ConverterDxt::decodeTile + decodeBlockDxt1/3/5 produce (tileconv/converter_dxt.cpp:147-176,251-430).
 * reference_bc3_alpha != 0 reproduces the reference's BC3 alpha decode, which reads its endpoints and
 * index bits from bytes 2..9 of the block (converter_dxt.cpp:351-359); 0 follows the BC3 format.
 * Quantising the pixels back to a palette (Colors::ARGBToPal, libimagequant) stays on the host. */
int tcb200_decode(tcb200_ctx* ctx, const uint8_t* records, int n, uint8_t* bgra, int reference_bc3_alpha);

/* Kernel-only timing aid for the decode kernel on device-resident records / pixels (same conventions as
 * tcb200_time_kernel_ms below); mean milliseconds per launch, negative on error. */
float tcb200_time_decode_ms(tcb200_ctx* ctx, const void* d_records, int n, void* d_bgra, int reference_bc3_alpha, int reps);

/* Kernel-only timing aid for bench.py: runs the encode kernel `reps` times on device-resident data
 * on the context's compute stream and returns the mean milliseconds per launch measured with CUDA
 * events on that stream (negative error code on failure). */
float tcb200_time_kernel_ms(tcb200_ctx* ctx, const void* d_tiles, const void* d_wh, int n, void* d_out, int reps);

/* Measured FP32 pipe peak of the device, for the ALU roofline: a register-resident loop of
 * independent fused multiply-adds (fused != 0) or of separate multiply and add instructions
 * (fused == 0, the form the bit-exact fitters are restricted to).  Returns lane-operations per
 * second counting one per FMUL/FADD/FFMA instruction lane (so FLOP/s = 2x for fused). */
double tcb200_fp32_peak(tcb200_ctx* ctx, int fused);

/* Device self-test of the kernels' inlined reciprocal (MUFU.RCP + one Newton step, used for the 2x2
 * least-squares denominators): compares it with IEEE 1.0f/x for every binary32 value of magnitude
 * 2^-44 .. 2^10, both signs, and +0.  Returns the number of mismatches (0 = bit-exact), negative on error. */
long long tcb200_selftest_reciprocal(tcb200_ctx* ctx);

/* Device self-test of the TCB200_RECIP_SSE estimate: estimates[i] = the value the kernels use in place of
 * the host's rcpps(values[i]) (both host arrays of n floats).  The context must have been created with
 * TCB200_RECIP_SSE.  The tests compare the result with the CPU instruction itself. */
int tcb200_selftest_rcpps(tcb200_ctx* ctx, const float* values, int n, float* estimates);

/* Host-side self-test of the candidate tables of the cluster sweeps (no device needed): every list holds exactly the
 * partitions libsquish's loops visit, in their order (3-colour: i, then j from max(i,1); 4-colour: i, j from i, k from
 * max(j,1)); every entry's four 16-bit fields are 4 x a prefix-table index <= 152, so that bits 14-15 of a low field
 * are zero (the kernels read the high field as word >> 14) and the bound-table offsets stay inside the table; and
 * every list is followed by at least 128 readable entries (the pruned sweep reads ahead).  Returns 0, or the number
 * of violations. */
int tcb200_selftest_tables(void);

/* Kernels launched by this context so far (bench.py's gpu_launches). */
long long tcb200_launch_count(const tcb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TCB200_H */
