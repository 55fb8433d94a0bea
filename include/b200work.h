/* b200work.h — device-side generators of the benchmark workloads + result checksums (exported by libb200aabb.so).
 *
 * NOT part of the reference's API: these entry points exist so that BASELINE.json's config 5 ("1 B mixed queries generated on device in
 * 64 M-query batches, results reduced to a 128-bit checksum per batch") runs without PCIe traffic in the timed region.
 * The generators reproduce cgal_b200/workloads.py (numpy, the oracle's input) bit for bit: a counter-based SplitMix64 of
 * (seed, stream, index), u = (x >> 11) * 2^-53, and the same binary64 operation order (tests/test_gpu_workload.py compares them).
 * All pointers are DEVICE pointers; work is enqueued on `cuda_stream` of `device` and NOT synchronised, except b200work_checksum128.
 * Every function returns 0 or a negative B200AABB_ERR_* code.
 */
#ifndef B200WORK_H
#define B200WORK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* workloads.bbox_points(bbox, n, seed, offset): out[3 i + k] = lo[k] + u(3 (offset + i) + k) * (hi[k] - lo[k]); stream 1 of `seed`. */
int b200work_bbox_points(int device, void* cuda_stream, const double bbox[6], uint32_t seed, uint64_t offset, uint64_t n, double* out);

/* workloads.pinhole_rays: rows [row0, row0 + nrows) of a width x height image, ray record = (eye, fl(eye + dir)).
 * camera[13] = eye(3), fwd(3), right(3), up(3), th = tan(vfov / 2) as the host computed them (numpy); out: nrows * width * 6 doubles. */
int b200work_pinhole_rays(int device, void* cuda_stream, const double camera[13], uint32_t width, uint32_t height, uint32_t row0, uint32_t nrows, double* out);

/* Order-independent 128-bit checksum of a device buffer read as little-endian 64-bit words (the tail is zero-padded): word i (global
 * position index0 + i) contributes mix64(w ^ K1 (pos + 1)) to sum[0] and mix64(w + K2 (pos + 1)) to sum[1], both modulo 2^64, so the
 * checksums of the shards of a batch add up to the checksum of the whole batch.  Synchronises the stream; sum is a HOST pointer. */
int b200work_checksum128(int device, void* cuda_stream, const void* data, uint64_t nbytes, uint64_t index0, uint64_t sum[2]);

/* FP64 pipe microbenchmark (8 independent DFMA chains per thread, 148 x 8 blocks): the measured peak in TFLOP/s (2 flops per DFMA). */
int b200work_dfma_tflops(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* B200WORK_H */
