/*
 * smap_b200 - debug / bisection entry points of libsmap_b200.so.  NOT part of the drop-in boundary (include/smap_b200.h);
 * used by tools/debug_*.py to localise numerical differences op by op.  They synchronise the device.
 */
#ifndef SMAP_B200_DEBUG_H
#define SMAP_B200_DEBUG_H

#include "smap_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* 64-bit position-weighted checksums of the output tensor of every op of the batch-B plan, as left by the last forward.
 * sums[max_ops]; desc (optional): max_ops strings of desc_stride bytes describing each op.  Returns the number of ops. */
int smapb_debug_checksums(smapb_handle* h, int B, unsigned long long* sums, int max_ops, char* desc, int desc_stride);
/* Raw copy (both bf16 planes, or fp32 for head outputs) of op `idx`'s output into host memory; returns the bytes copied. */
long long smapb_debug_dump(smapb_handle* h, int B, int idx, void* host, long long max_bytes, int which);

/* Host-only: the resampling plan smapb_preprocess uses for a src_w x src_h image (no GPU work).  dims6 = {dst_w, dst_h,
 * pad_left, pad_top, mode (0 fixed-point bilinear, 1 exact 1/2 scale = 2x2 rounded mean, 2 copy), 0}; the tables (may be
 * NULL) must hold dst_w, 2*dst_w, 2*dst_h and 2*dst_h entries (dst <= net size). */
int smapb_debug_resize_plan(int src_w, int src_h, int net_w, int net_h, int* dims6, double* scale, int* xofs, short* xcoef,
                            int* yofs, short* ycoef);

/* Environment switches read when a handle / plan is built (never on the per-call path):
 *   SMAPB_DEBUG_STOP=n        run only the first n ops of the plan
 *   SMAPB_DEBUG_SYNC=1        synchronise the stream after every launch
 *   SMAPB_DEBUG_ONEGROUP=fpro one epilogue group for fp32-out / post-add / residual / other layers
 *   SMAPB_FORCE_TILE=bn,cg    force a tile shape wherever it is valid;  SMAPB_NO_AUTOTUNE=1  cost model only
 *   SMAPB_PAIR=0|1|2          CTA pairs off / model / always;  SMAPB_NO_BN256=1  no one-CTA 128x256 tiles;  SMAPB_NO_HALO=1  no halo strips
 *   SMAPB_ONE_STREAM=1        no side stream;  SMAPB_NO_GRAPH=1  no CUDA graph replay;  SMAPB_PDL=1  programmatic dependent launch
 *   SMAPB_STEM=cuda           CUDA-core stem;  SMAPB_NO_FUSE_DS=1 / SMAPB_NO_FUSE_UP=1  unfused downsample / up-residual
 *   SMAPB_ROLES=1             per-role wait-cycle counters in smapb_conv_test
 *   SMAPB_ROLES_PLAN=file.csv the same counters for every conv launch of a profiled (smapb_profile_begin/end) run, i.e. inside the
 *                             real step (tools/roles_plan.py)
 *   SMAPB_TIMELINE=1          clock64 time line of CTA 0 of one launch in smapb_conv_test (set-up, first operands, main loop end,
 *                             chunk ends, exit)
 *   SMAPB_LIB=path (Python)   load another build of the library (A/B runs: tools/gpu_ab.sh, tools/ab_hash.py) */

#ifdef __cplusplus
}
#endif
#endif /* SMAP_B200_DEBUG_H */
