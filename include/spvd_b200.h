/*
 * spvd_b200.h -- C ABI of libspvd_b200.so: the sm_100a (B200) backend for the hot path of the SPVD
 * sparse point-voxel denoiser (JohnRomanelis/SPVD).
 *
 * This is the drop-in boundary.  The reference obtains these operations from un-vendored Python
 * packages (torchsparse 2.1.0, torch-scatter 2.1.2); every entry point below names the reference
 * call site (file:line under the reference checkout) whose backend work it replaces.  A
 * reference-side binding is a ctypes stub per function; see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host"; nothing here allocates or frees;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point
 *     synchronises the device, so calls are CUDA-graph capturable;
 *   - the return value is SPVD_OK or an SPVD_ERR_* code; spvd_last_error() gives the text
 *     (thread-local).  Device-side anomalies (coordinate outside the packed range, hash table
 *     full) are OR-ed into the optional `status` word, which the caller reads whenever it next
 *     synchronises;
 *   - coordinates are int32 [N,4] = (batch, x, y, z) (reference: models/sparse_utils.py:44,62,83);
 *     -32767 <= x,y,z <= 32767 and 0 <= batch <= 65534;
 *   - row counts: `n_cap` is the host-known number of rows (or an upper bound); when the
 *     optional `n_dev` (device int32) is non-NULL the kernels use min(*n_dev, n_cap) rows, which
 *     lets a caller keep true sizes on the device;
 *   - kernel maps are stored offset-major: map_t[k*ld + o] = input row feeding output row o
 *     through kernel offset k, or -1; ld >= n_out, a multiple of 128; rows o >= n_out hold -1.
 *     tile_mask[o/128] has bit k set when any of the tile's 128 output rows uses offset k.
 */
#ifndef SPVD_B200_H
#define SPVD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPVD_ABI_VERSION 1

#define SPVD_OK 0
#define SPVD_ERR_INVALID 1   /* bad argument */
#define SPVD_ERR_CUDA 2      /* a CUDA runtime call or launch failed */
#define SPVD_ERR_WORKSPACE 3 /* workspace too small */

#define SPVD_STATUS_COORD_RANGE 1 /* a coordinate did not fit the 16-bit packed key */
#define SPVD_STATUS_HASH_FULL 2   /* insertion found no free slot */
#define SPVD_STATUS_SEGMENT_OVERFLOW 4 /* a cloud has more rows than the max_cloud_rows hint */

#define SPVD_SCALE_MUL 0 /* fl(fl(c*init_res) * fl(1/after_res)): torch-CUDA's division by a scalar */
#define SPVD_SCALE_DIV 1 /* fl(fl(c*init_res) / after_res):       torch-CPU's                        */

#define SPVD_DOWNSAMPLE_SPCONV 0 /* torchsparse 2.1 default: drop parents above (max_in-1)//2 */
#define SPVD_DOWNSAMPLE_FLOOR 1  /* keep every floor(c/2) */

#define SPVD_ORDER_BXYZ 0 /* coordinate columns (batch,x,y,z) */
#define SPVD_ORDER_XYZB 1 /* coordinate columns (x,y,z,batch): what sphashquery passes, models/sparse_utils.py:44,51 */

#define SPVD_CONV_AUTO 0
#define SPVD_CONV_SIMT 1 /* fp32 FMA reference kernel (any channel count) */
#define SPVD_CONV_UMMA 2 /* tcgen05 implicit GEMM, TF32 operands, fp32 accumulate in TMEM */

typedef void* spvd_stream_t;

const char* spvd_last_error(void);
int spvd_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * V1  initial_voxelize, coordinate half           reference: models/sparse_utils.py:60-66
 * pc      [n,4] fp32 = (batch, x, y, z) point coordinates in units of init_res (x.C.float(),
 *         models/spvd.py:436)
 * fcoords [n,4] fp32 = (batch, scaled x,y,z)      -> becomes z.C (models/sparse_utils.py:72)
 * icoords [n,4] int32 = floor(fcoords)            -> new_int_coord (models/sparse_utils.py:65)
 */
int spvd_point_voxel_coords(const float* pc, int n, float init_res, float after_res, int scale_mode,
                            float* fcoords, int32_t* icoords, spvd_stream_t stream);

/* torch.unique(int32[N,4], dim=0)                reference: models/sparse_utils.py:66
 * out_coords [n_cap,4] sorted lexicographically by (batch,x,y,z), first *out_count rows valid.  The
 * inverse index (idx_query of initial_voxelize, models/sparse_utils.py:67) is a kernel_size-1
 * spvd_hash_query of the points against a table built on out_coords, exactly as the reference
 * obtains it from sphashquery. */
size_t spvd_unique_workspace_bytes(int n_cap);
int spvd_unique_coords(const int32_t* icoords, int n_cap, const int32_t* n_dev, int32_t* out_coords,
                       int32_t* out_count, int32_t* status, void* workspace, size_t workspace_bytes,
                       spvd_stream_t stream);

/* Segmented variants of spvd_unique_coords (parent = 0) / spvd_downsample_coords (parent = 1) for
 * batch-major rows with at most max_cloud_rows (<= 16384) rows per cloud: every cloud is sorted and
 * de-duplicated by one CTA in shared memory (3 launches instead of ~20).  Also writes the new level's
 * per-cloud counts [n_batch] and offsets [n_batch+1].  A cloud larger than the hint sets
 * SPVD_STATUS_SEGMENT_OVERFLOW.  Workspace: spvd_unique_workspace_bytes(n_cap). */
int spvd_segmented_unique(const int32_t* coords, int n_cap, const int32_t* n_dev, int n_batch,
                          int max_cloud_rows, int parent, int downsample_mode, int32_t* out_coords,
                          int32_t* out_count, int32_t* batch_counts, int32_t* batch_offsets, int32_t* status,
                          void* workspace, size_t workspace_bytes, spvd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * H1  coordinate hash table                      reference: models/sparse_utils.py:36-55
 * (torchsparse.backend.GPUHashTable: keys int64[capacity] and vals int32[capacity] are the
 * caller's zero-initialised tensors, models/sparse_utils.py:37-42).  Values are row+1, 0 = absent,
 * exactly what lookup_coords returns before the reference subtracts 1 (models/sparse_utils.py:49-54).
 * kernel_size 1, 2 ({0,1}^3, z fastest) or 3 ({-1,0,1}^3, x fastest); out is [m, kernel_size^3].
 */
int spvd_hash_build(const int32_t* coords, int n_cap, const int32_t* n_dev, int coord_order,
                    int64_t* keys, int32_t* vals, int capacity, int32_t* status,
                    spvd_stream_t stream);
int spvd_hash_query(const int32_t* query, int m, int coord_order, int kernel_size,
                    const int64_t* keys, const int32_t* vals, int capacity, int32_t* out,
                    spvd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K1  submanifold kernel map, kernel 3 stride 1  reference: spnn.Conv3d(k=3) at models/spvd.py:35,38,
 *     models/ddpm_unet_attn.py:28 (torchsparse build_kernel_map, frames at
 *     experiments/ConditionalGeneration.ipynb:449-451)
 * map_t [27, ld], tile_mask [ld/128].
 */
int spvd_kmap_subm3(const int32_t* coords, int n_cap, const int32_t* n_dev, const int64_t* keys,
                    const int32_t* vals, int capacity, int32_t* map_t, int ld, uint32_t* tile_mask,
                    spvd_stream_t stream);

/* K2  stride-2 kernel-2 output coordinates       reference: spnn.Conv3d(ni,nf,2,stride=2) at
 *     models/spvd.py:113 (torchsparse spdownsample; boundary rule: SURVEY.md Appendix B)
 * out_coords [n_cap,4] = unique floor(c/2), sorted, first *out_count rows valid. */
size_t spvd_downsample_workspace_bytes(int n_cap);
int spvd_downsample_coords(const int32_t* coords, int n_cap, const int32_t* n_dev, int mode,
                           int32_t* out_coords, int32_t* out_count, int32_t* status, void* workspace,
                           size_t workspace_bytes, spvd_stream_t stream);
/* down map [8, ld_coarse] (coarse <- fine children, offset = child parity 4*px+2*py+pz) and K3 its
 * transpose [8, ld_fine] (fine <- parent; reference: spnn.Conv3d(..., transposed=True) at
 * models/spvd.py:229), from a hash table built on the coarse coordinates.  Any output may be NULL.
 * Both maps are fully written (absent = -1) together with their tile masks; parent [n_fine_cap] =
 * coarse row of each fine voxel or -1 when its parent was dropped, koff [n_fine_cap] its offset. */
int spvd_kmap_down2(const int32_t* fine_coords, int n_fine_cap, const int32_t* n_fine_dev,
                    const int64_t* keys, const int32_t* vals, int capacity, int32_t* map_down_t,
                    int ld_coarse, uint32_t* tile_mask_down, int32_t* map_up_t, int ld_fine,
                    uint32_t* tile_mask_up, int32_t* parent, int32_t* koff, spvd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * V2  point -> voxel index at tensor stride s    reference: models/sparse_utils.py:79-90
 * pc [n,4] fp32 = z.C after initial_voxelize (batch, stride-1 voxel coords as floats);
 * idx [n] = row of voxel (int(batch), floor(xyz / s)) or -1 (the caller clamps, :93). */
int spvd_point_voxel_query(const float* pc, int n, int stride, const int64_t* keys,
                           const int32_t* vals, int capacity, int32_t* idx, spvd_stream_t stream);

/* V3  trilinear corner indices + weights          reference: models/sparse_utils.py:108-114
 * (sphashquery(kernel_size=2) + torchsparse calc_ti_weights(scale=1)); idx8/w8 are [n,8]. */
int spvd_devox_query(const float* pc, int n, int stride, const int64_t* keys, const int32_t* vals,
                     int capacity, int32_t* idx8, float* w8, spvd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * torch_scatter.scatter_mean(src, idx, dim=0)     reference: models/sparse_utils.py:69,94
 * out [nv,c] and counts [nv] are written (zero-filled inside); backward: gsrc = gout[idx]/max(cnt,1).
 */
int spvd_scatter_mean_fwd(const float* src, const int32_t* idx, int n, int c, float* out,
                          float* counts, int nv, spvd_stream_t stream);
int spvd_scatter_mean_bwd(const float* gout, const int32_t* idx, const float* counts, int n, int c,
                          float* gsrc, spvd_stream_t stream);
/* The same mean with the weight 1 / (points in the point's voxel) precomputed once per level (spvd_point_mean_weights;
 * counts = int32 scratch [nv_cap]): spvd_scatter_wmean_fwd is then one zero-fill and one RED pass -- no count accumulation,
 * no divide pass over the voxel rows.  Differs from spvd_scatter_mean_fwd by the rounding of x/n vs x*(1/n) only. */
int spvd_point_mean_weights(const int32_t* idx, int n, int nv_cap, int32_t* counts, float* w, spvd_stream_t stream);
int spvd_scatter_wmean_fwd(const float* src, const int32_t* idx, const float* w, int n, int c, float* out, int nv,
                           spvd_stream_t stream);
/* The same mean as a segmented reduction, for levels where many points share a voxel (same-address RED serialises: at
 * stride 16 a voxel holds ~35 points and row 0, the sink of unmatched points, thousands).  spvd_group_points_by_voxel (once
 * per level): counts [nv_cap + 2], start / cursor [nv_cap], order [n] = point ids voxel by voxel (voxel v owns
 * order[start[v] .. start[v] + counts[v])), units [2 * (nv_cap + n / 32 + 1)] = (voxel, chunk) records, one per <= 32 points
 * of a voxel.  spvd_segment_mean_fwd: zero-fill, then one thread per (unit, 4 channels) sums the unit's rows and adds its
 * share of the mean with one vector RED.  c % 4 == 0. */
int spvd_group_points_by_voxel(const int32_t* idx, int n, int nv_cap, const int32_t* nv_dev, int32_t* counts,
                               int32_t* start, int32_t* cursor, int32_t* order, int32_t* units, spvd_stream_t stream);
int spvd_segment_mean_fwd(const float* src, const int32_t* start, const int32_t* counts, const int32_t* order,
                          const int32_t* units, int n, int nv_cap, int c, float* out, int nv, spvd_stream_t stream);

/* torchsparse.nn.functional.spdevoxelize          reference: models/sparse_utils.py:119,128
 * out[p] = sum_j w8[p,j] * feats[idx8[p,j]] (idx -1 skipped); backward scatter-adds into gfeats
 * [nv,c] (zero-filled inside). */
int spvd_devoxelize_fwd(const float* feats, const int32_t* idx8, const float* w8, int n, int c,
                        float* out, spvd_stream_t stream);
int spvd_devoxelize_bwd(const float* gout, const int32_t* idx8, const float* w8, int n, int c,
                        float* gfeats, int nv, spvd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * C1-C4  sparse convolution  out[o] = sum_k in[map_t[k][o]] @ W[k] (+ bias)
 *     reference: spnn.Conv3d.forward at models/spvd.py:35,38,113,229; models/ddpm_unet_attn.py:28
 * w        [kvol,cin,cout] fp32, the reference's `kernel` parameter layout
 * w_packed the same weights in the tensor-core tile image produced by spvd_conv_pack_weights
 *          (may be NULL when impl == SPVD_CONV_SIMT)
 * map_t    [kvol, ld] or NULL for the identity map (kernel_size 1)
 * The UMMA path needs cin % 32 == 0 and cout % 32 == 0 (operands are rounded to TF32, round-to-nearest,
 * inside the kernel / the packer); SPVD_CONV_AUTO picks it whenever those hold and w_packed is given.
 */
int spvd_round_tf32(const float* x, float* y, long long count, spvd_stream_t stream);
/* w_t[k'][co][ci] = w[k][ci][co] with k' = flip ? kvol-1-k : k  (weights of the data gradient) */
int spvd_conv_transpose_weights(const float* w, int kvol, int cin, int cout, int flip, float* w_t,
                                spvd_stream_t stream);
size_t spvd_conv_packed_weight_count(int kvol, int cin, int cout);
int spvd_conv_pack_weights(const float* w, int kvol, int cin, int cout, float* w_packed,
                           spvd_stream_t stream);
int spvd_conv_umma_supported(int kvol, int cin, int cout);
int spvd_conv_fwd(const float* in, const float* w, const float* w_packed, const int32_t* map_t, int ld,
                  const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout,
                  const float* bias, float* out, int impl, spvd_stream_t stream);
/* residual (optional, [n_out,cout], must not alias out): out = conv + bias + residual in the epilogue
 * (the skip connection of EmbResBlock, models/ddpm_unet_attn.py:60, and of PointBlock, models/spvd.py:27) */
int spvd_conv_fwd_simt(const float* in, const float* w, const int32_t* map_t, int ld,
                       const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout,
                       const float* bias, const float* residual, float* out, spvd_stream_t stream);
/* bn = output-channel slice width per CTA (32,64,96,128,192,256 dividing cout; 0 = widest);
 * nsplit = number of CTAs sharing one tile's (offset, 32-channel slab) iterations, partial sums
 * accumulated with vector RED into the zero-filled output (0 = choose from the grid size);
 * a_is_tf32 != 0 promises that `in` is already TF32-representable (spvd_round_tf32, or a producer
 * that rounds, e.g. spvd_bn_silu_fwd with act & 2): the gather then runs as cp.async with kStages
 * slabs in flight instead of through registers with cvt.rna;
 * mt = tile configuration: 0 / 3 = one 128-row tile per CTA, two CTAs per SM (the product path); 1 = one CTA per SM with a
 * four-stage ring.  (Cluster multicast of the weight tile and two tiles per CTA exist behind -DSPVD_EXPERIMENTAL: correct,
 * measured no faster; the round-1 stream-K and CTA-pair kernels were removed)
 *
 * f16 operands (a_is_tf32 = SPVD_A_F32_TO_F16 or SPVD_A_F16): same tiles with A and B in IEEE half -- an 11-bit
 * significand, i.e. TF32's precision, fp32 accumulation -- on tcgen05.mma.kind::f16: twice the TF32 issue rate and 64
 * input channels per 128-byte slab, so half the iterations and half the operand bytes.  w_packed must then be the
 * spvd_conv_pack_weights_f16 image; `in` is fp32 (converted on the fly, clamped to the finite half range) for
 * SPVD_A_F32_TO_F16, or an f16 [rows, cin] tensor (a producer with act & 4) for SPVD_A_F16. */
#define SPVD_A_F32 0        /* fp32 rows, rounded to TF32 while gathered          */
#define SPVD_A_TF32 1       /* fp32 rows that are already TF32-representable      */
#define SPVD_A_F32_TO_F16 2 /* fp32 rows, converted to f16 while gathered         */
#define SPVD_A_F16 3        /* f16 rows                                           */
size_t spvd_conv_packed_weight_count_f16(int kvol, int cin, int cout); /* in halves */
int spvd_conv_pack_weights_f16(const float* w, int kvol, int cin, int cout, void* w_packed, spvd_stream_t stream);
int spvd_conv_fwd_umma(const float* in, const float* w_packed, const int32_t* map_t, int ld,
                       const uint32_t* tile_mask, int n_out, int kvol, int cin, int cout,
                       const float* bias, const float* residual, float* out, int bn, int nsplit,
                       int a_is_tf32, int mt, spvd_stream_t stream);
/* Same kernels on mask-sorted tiles (spvd_kmap_sort): map_t / tile_mask are in the sorted row order and tile row j
 * writes output row out_rows[j] (a permutation of 0..n_out-1; NULL = identity, i.e. spvd_conv_fwd_umma).  Rows that share
 * their set of active kernel offsets sit in the same 128-row tile, so far fewer (tile, offset) steps multiply zero rows --
 * what torchsparse calls a "sorted implicit GEMM"; replaces the conv launch inside spnn.Conv3d (models/spvd.py:35,38).
 * mt = 4 with SPVD_A_F16 / SPVD_A_TF32: clusters of two CTAs multicast each half of every weight tile to both. */
int spvd_conv_fwd_umma_rows(const float* in, const float* w_packed, const int32_t* map_t, int ld,
                            const uint32_t* tile_mask, const int32_t* out_rows, int n_out, int kvol, int cin, int cout,
                            const float* bias, const float* residual, float* out, int bn, int nsplit,
                            int a_is_tf32, int mt, spvd_stream_t stream);
/* Mask-sorted view of a kernel map: rows sorted by the bit set of their active offsets (stable: ties keep row order).
 * map_sorted [kvol,ld] = the map with its columns in that order, mask_sorted [ld/128] = per-tile OR of the row masks,
 * perm [ld] = sorted position -> original row (-1 past the row count).  workspace >= spvd_kmap_sort_workspace_bytes(n_cap). */
size_t spvd_kmap_sort_workspace_bytes(int n_cap);
int spvd_kmap_sort(const int32_t* map_t, int ld, int n_cap, const int32_t* n_dev, int kvol, int32_t* map_sorted,
                   uint32_t* mask_sorted, int32_t* perm, void* workspace, size_t workspace_bytes, spvd_stream_t stream);
/* Persistent sorted-tile convolution with a fused epilogue (csrc/conv_sp_f16.cu), f16 operands, fp32 accumulate:
 *   v[o] = sum_k in[map_t[k][j]] W[k] (+ in2[o] W2) (+ bias) (+ residual[o]);  FiLM: v = (1 + film[b(o)][c]) v + film[b(o)][cout + c]
 *   out32[o] = v;   out16[o] = f16(act(v * aff_scale + aff_shift))     (o = out_rows[j], or j when out_rows is NULL)
 * in / in2 are f16 row-major, w / w2 spvd_conv_pack_weights_f16 images ([kvol,cin,cout] and [1,cin2,cout]); map_t NULL =
 * identity (kvol 1).  Replaces spnn.Conv3d + the FiLM / BatchNorm / SiLU / shortcut glue of EmbResBlock.forward
 * (models/ddpm_unet_attn.py:48-65).  bn = output slice width (0 = choose), stages = stage-ring cap (0 = choose). */
typedef struct spvd_conv_sp_args {
  const void* in;
  const void* w;
  const int32_t* map_t;
  const uint32_t* tile_mask;
  const int32_t* out_rows;
  const void* in2;
  const void* w2;
  const float* bias;
  const float* residual;
  const float* film;     /* [n_batch, film_ld]: scale at column c, shift at column cout + c */
  const int32_t* coords; /* [n_out,4], column 0 = cloud index (FiLM only) */
  const float* aff_scale;
  const float* aff_shift;
  float* out32;
  void* out16;
  int32_t ld, n_out, kvol, cin, cout, cin2, film_ld, act; /* act bit 0 = SiLU on the f16 output */
  int32_t n_in, pad;                                       /* rows of `in` (bound of the gather; 0 = unknown) */
} spvd_conv_sp_args_t;
int spvd_conv_sp_f16(const spvd_conv_sp_args_t* args, int bn, int stages, spvd_stream_t stream);
/* Pair-major view of a kernel map, for sparse levels (few active offsets per voxel): per offset the valid
 * (input row, output row) pairs compacted into tiles of 128 pairs, ordered by offset then output row.
 * pair_in / pair_out [tile_cap*128] (-1 = padding), tile_k [tile_cap] = offset of each tile,
 * *n_tiles / *n_pairs device counts, scratch >= 32 + kvol * ceil(n_cap / 4096) int32. */
int spvd_kmap_pairs(const int32_t* map_t, int ld, int n_cap, const int32_t* n_dev, int kvol, int tile_cap,
                    int32_t* pair_in, int32_t* pair_out, int32_t* tile_k, int32_t* n_tiles, int32_t* n_pairs,
                    int32_t* scratch, spvd_stream_t stream);
/* Same convolution, pair-major: one CTA per pair tile, K loop over the channel slabs only, product rows
 * scattered into the zero-filled output with vector RED; bias / residual are added by the tiles of the
 * centre offset (center_k; pass -1 for maps without one -- the stride-2 maps -- which then take neither).
 * Swapping pair_in and pair_out gives the transposed convolution on the same lists. */
int spvd_conv_fwd_pairs(const float* in, const float* w_packed, const int32_t* pair_in,
                        const int32_t* pair_out, const int32_t* tile_k, int n_pair_tiles, int center_k,
                        int n_out, int kvol, int cin, int cout, const float* bias, const float* residual,
                        float* out, int a_is_tf32, spvd_stream_t stream);
/* The fp32 SIMT version for the network input (cin <= 4, e.g. the 3 -> 32 stem convolution): one warp instruction per
 * pair, RED into the zero-filled output, bias added by the centre-offset tiles. */
int spvd_conv_fwd_pairs_simt(const float* in, const float* w, const int32_t* pair_in, const int32_t* pair_out,
                             const int32_t* tile_k, int n_pair_tiles, int center_k, int n_out, int kvol, int cin,
                             int cout, const float* bias, float* out, spvd_stream_t stream);
/* Weight gradient on the tensor cores (tcgen05, TF32 operands -- round x and gout first with spvd_round_tf32 --,
 * fp32 accumulate): both operands are gathered rows fed as MN-major SWIZZLE_128B tiles, the reduction runs over
 * the pair lists of spvd_kmap_pairs (pair_in = rows of x, pair_out = rows of gout; all NULL with kvol == 1 =
 * identity pairs over n_identity rows).  gw [kvol,cin,cout] is zero-filled inside; partial sums are RED-added. */
int spvd_conv_wgrad_umma(const float* x, const float* gout, const int32_t* pair_in, const int32_t* pair_out,
                         const int32_t* tile_k, int n_pair_tiles, int n_identity, int kvol, int cin, int cout,
                         float* gw, spvd_stream_t stream);
/* gw[k] = sum over pairs (i,o) of offset k of in[i]^T gout[o];  gw [kvol,cin,cout] zero-filled inside */
int spvd_conv_wgrad(const float* in, const float* gout, const int32_t* map_t, int ld, int n_out,
                    int kvol, int cin, int cout, float* gw, spvd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * C5/M1  inference-time glue fused into single passes
 * spvd_bn_silu_fwd: y = act(x*scale[c] + shift[c]) -- spnn.BatchNorm (eval: scale = w/sqrt(var+eps),
 *     shift = b - mean*scale) followed by spnn.SiLU (act & 1) or nothing; act & 2 additionally rounds
 *     y to TF32 (round-to-nearest) so that the following convolution can gather it with cp.async;
 *     reference: models/ddpm_unet_attn.py:24-28, models/spvd.py:36-37.
 * spvd_film_fwd: y[i] = (1 + tt[b_i, c]) * x[i, c] + tt[b_i, C + c], b_i = coords[i].batch, tt [B, 2C];
 *     reference: TimeEmbeddingBlock.message, models/modules/graph.py:35-41.
 * spvd_batch_counts: counts[b] = number of rows of batch b; rows must be batch-major (every voxel level
 *     of this library is; what to_dense_batch derives,
 *     models/modules/transformer.py:12-15). */
int spvd_bn_silu_fwd(const float* x, const float* scale, const float* shift, long long rows, int c,
                     int act, float* y, spvd_stream_t stream);
int spvd_film_fwd(const float* x, const float* tt, const int32_t* coords, long long rows, int c,
                  float* y, spvd_stream_t stream);
int spvd_batch_counts(const int32_t* coords, int n_cap, const int32_t* n_dev, int nb, int32_t* counts,
                      spvd_stream_t stream);

/* FiLM + the BatchNorm/SiLU that follows it in EmbResBlock (models/ddpm_unet_attn.py:54-57) in one
 * pass: y = act(((1 + tt[b,c]) * x + tt[b,C+c]) * scale[c] + shift[c]); tt rows are ldt floats apart. */
int spvd_film_affine_fwd(const float* x, const float* tt, int ldt, const int32_t* coords,
                         const float* scale, const float* shift, long long rows, int c, int act,
                         float* y, spvd_stream_t stream);
/* Timestep embedding of the denoiser in one launch (models/utils.py:15-19 sinusoid; models/spvd.py:380-392 emb_mlp =
 * BatchNorm1d(eval) -> SiLU -> Linear(n_temb, n_emb) -> SiLU -> Linear(n_emb, n_emb); models/spvd.py cond_emb adds a
 * class row): out[b] = act(W2 silu(W1 silu(bn(sincos(t[b] * freqs))) + b1) + b2 + cond_table[cls[b]]).
 * params = freqs [n_temb/2] | bn scale [n_temb] | bn shift [n_temb] | W1 [n_emb,n_temb] | b1 | W2 [n_emb,n_emb] | b2;
 * cls / cond_table may be NULL; act & 1 applies the SiLU every TimeEmbeddingBlock starts with. */
int spvd_time_embed_fwd(const int64_t* t, const int64_t* cls, const float* params, const float* cond_table,
                        int n_batch, int n_temb, int n_emb, int act, float* out, spvd_stream_t stream);
/* torchsparse.cat (models/spvd.py:232) fused with the BatchNorm+SiLU that follows it in the up-path EmbResBlock
 * (models/ddpm_unet_attn.py:24-28): cat_f16 = [a | b] and y_f16 = act([a | b] * scale + shift), both stored as f16 -- the
 * operands of the block's 1x1 shortcut and of its first 3^3 convolution. */
int spvd_cat_affine_f16_fwd(const float* a, const float* b, long long rows, int ca, int cb, const float* scale,
                            const float* shift, int act, void* cat_f16, void* y_f16, spvd_stream_t stream);
/* torchsparse.cat (models/spvd.py:232): y = [a | b] along channels */
int spvd_cat_fwd(const float* a, const float* b, long long rows, int ca, int cb, float* y,
                 spvd_stream_t stream);
/* SparseAttention (models/modules/transformer.py:102-125): LayerNorm, then per-cloud multi-head
 * attention over batch-major rows (batch_offsets [n_batch+1] = first row of each cloud); qkv is
 * [rows, 3c] laid out which*c + head*d + d_i.  act & 2 rounds the output to TF32. */
int spvd_layernorm_fwd(const float* x, const float* gamma, const float* beta, int rows, int c,
                       float eps, int act, float* y, spvd_stream_t stream);
int spvd_attention_fwd(const float* qkv, const int32_t* batch_offsets, int n_batch, int max_len, int c,
                       int heads, int act, float* out, spvd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * N1  per-step work of the sampling loop around the network, fused:
 *   x_new = c1 * (x_t - c2 * eps) + sigma * z        DDPM.update_rule, utils/schedulers.py:78-89 (z NULL = last step)
 *   coords = (cloud, floor((x_new - min_cloud x_new) / pres))   SparseSchedulerGPU.torch2sparse, utils/schedulers.py:249-257
 * x_t / eps / z / x_new are [n_clouds*n_per_cloud, f] (xyz first); eps NULL = only re-sparsify x_t.  coords int32 [.,4] and/or
 * pc float [.,4] (what SPVUnet.forward feeds the planner, x.C.float()) may be NULL; scratch >= 3*n_clouds int32. */
int spvd_ddpm_step(const float* x_t, const float* eps, const float* z, float c1, float c2, float sigma, int n_clouds,
                   int n_per_cloud, int f, float pres, float* x_new, int32_t* coords, float* pc, int32_t* scratch,
                   spvd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Geometry planner: V1 + K1 + K2/K3 + V2/V3 for every tensor stride of the U-Net in one call, into
 * caller-owned buffers (level i has tensor stride 2^i; NULL pointers = not needed at that level).
 * Everything the reference caches lazily in _caches (models/sparse_utils.py:71,90,122-125 and the
 * torchsparse kmaps/cmaps).  All buffers have n_points rows of capacity; maps have leading
 * dimension ld.  When summary_host is non-NULL the call ends with ONE device->host copy of
 * summary_count int32 from summary_dev (the caller lays the level counts, the status word and the
 * per-cloud counts out contiguously there) followed by a stream synchronise -- the only host sync
 * of a forward pass. */
typedef struct spvd_plan_level {
  int32_t* coords;  /* [n_points,4] */
  int32_t* count;   /* [1] */
  int64_t* keys;    /* hash table, capacity >= 2*n_points */
  int32_t* vals;
  int32_t* kmap3;   /* [27,ld] */
  uint32_t* mask3;  /* [ld/128] */
  int32_t* kmap_down; /* [8,ld]: next coarser level <- this level */
  uint32_t* mask_down;
  int32_t* kmap_up;   /* [8,ld]: this level <- next coarser level */
  uint32_t* mask_up;
  int32_t* p2v;       /* [n_points], misses clamped to row 0 (models/sparse_utils.py:93) */
  int32_t* devox_idx; /* [n_points,8] */
  float* devox_w;     /* [n_points,8] */
  int32_t* batch_counts;  /* [n_batch] */
  int32_t* batch_offsets; /* [n_batch+1] */
  int32_t* pair_in;       /* pair-major view of kmap3 (spvd_kmap_pairs), [pair_tile_cap*128] */
  int32_t* pair_out;
  int32_t* pair_tile_k;   /* [pair_tile_cap] */
  int32_t* pair_counts;   /* [2] = (tiles, pairs) */
  int32_t* pair_scratch;  /* >= 32 + 27*ceil(n_points/4096) int32 */
  int32_t* dpair_in;      /* pair-major view of kmap_down: fine rows */
  int32_t* dpair_out;     /*                               coarse rows */
  int32_t* dpair_tile_k;
  int32_t* dpair_counts;  /* [2] */
  float* p2v_w;           /* [n_points] 1 / (points sharing the point's voxel), with p2v; or NULL */
  int32_t* p2v_cnt;       /* [n_points + 2] points per voxel (+ 2 counters) */
  int32_t* p2v_start;     /* [n_points] with p2v_order / p2v_cursor / p2v_units: points grouped by voxel */
  int32_t* p2v_cursor;    /* [n_points] (spvd_group_points_by_voxel); NULL = not wanted for this level */
  int32_t* p2v_order;     /* [n_points] */
  int32_t* p2v_units;     /* [2 * (n_points + n_points / 32 + 1)] */
  int32_t* kmap3_sorted;      /* mask-sorted views (spvd_kmap_sort) of kmap3 / kmap_down / kmap_up: [kvol,ld] maps in sorted  */
  uint32_t* mask3_sorted;     /* row order, their per-tile offset masks [ld/128] and the row permutations [ld]; NULL = not   */
  int32_t* perm3;             /* wanted.  sort_ws (>= spvd_kmap_sort_workspace_bytes(n_points)) is this level's scratch.     */
  int32_t* kmap_down_sorted;
  uint32_t* mask_down_sorted;
  int32_t* perm_down;
  int32_t* kmap_up_sorted;
  uint32_t* mask_up_sorted;
  int32_t* perm_up;
  void* sort_ws;
  int32_t capacity, pair_tile_cap, dpair_tile_cap, sort_ws_bytes;
} spvd_plan_level_t;

/* max_cloud_rows > 0: the caller promises batch-major points with at most that many per cloud (the
 * layout torch2sparse produces) -> the segmented sort path; 0 = no assumption (global radix sort). */
int spvd_plan_build(const float* pc, int n_points, int n_batch, int max_cloud_rows, float init_res, float after_res,
                    int scale_mode, int downsample_mode, const spvd_plan_level_t* levels, int n_levels,
                    int ld, float* fcoords, int32_t* icoords, int32_t* status, const int32_t* summary_dev,
                    int32_t* summary_host, int summary_count, void* workspace, size_t workspace_bytes,
                    spvd_stream_t stream);

/* The same for the levels [level_begin, level_end) only (the levels below level_begin must have been planned into
 * the same buffers).  A forward can plan its finest levels, start the feature path on them, and plan the coarser ones
 * on a second stream meanwhile; sync = 0 leaves the final copy of the summary asynchronous on `stream`. */
int spvd_plan_build_range(const float* pc, int n_points, int n_batch, int max_cloud_rows, float init_res,
                          float after_res, int scale_mode, int downsample_mode, const spvd_plan_level_t* levels,
                          int n_levels, int level_begin, int level_end, int ld, float* fcoords, int32_t* icoords,
                          int32_t* status, const int32_t* summary_dev, int32_t* summary_host, int summary_count,
                          int sync, void* workspace, size_t workspace_bytes, spvd_stream_t stream);

/* Level groups [bounds[g], bounds[g+1]), g < n_groups, planned by one host call that never waits: group 0 on `stream`,
 * the others on `side_stream` behind it; events[g] (a cudaEvent_t) is recorded after group g's summary reached
 * summary_hosts[g].  The caller waits on events[0], starts the feature path on the finest levels, and collects the
 * coarser groups as their events complete. */
int spvd_plan_build_groups(const float* pc, int n_points, int n_batch, int max_cloud_rows, float init_res,
                           float after_res, int scale_mode, int downsample_mode, const spvd_plan_level_t* levels,
                           int n_levels, const int32_t* bounds, int n_groups, int ld, float* fcoords,
                           int32_t* icoords, int32_t* status, const int32_t* summary_dev,
                           int32_t* const* summary_hosts, int summary_count, void* const* events, void* workspace,
                           size_t workspace_bytes, spvd_stream_t stream, spvd_stream_t side_stream);

/* The same as CUDA graphs.  The planner's launch sequence depends only on capacities (grids are sized by n_points, row
 * counts stay on the device), so it is captured once per buffer set -- one graph per level group -- and replayed with one
 * launch per group.  pc_stage [n_points,4] is a buffer of the caller that launch() copies the points into. */
typedef struct spvd_plan_graph spvd_plan_graph_t;
int spvd_plan_graph_create(int n_points, int n_batch, int max_cloud_rows, float init_res, float after_res,
                           int scale_mode, int downsample_mode, const spvd_plan_level_t* levels, int n_levels,
                           const int32_t* bounds, int n_groups, int ld, float* pc_stage, float* fcoords,
                           int32_t* icoords, int32_t* status, const int32_t* summary_dev,
                           int32_t* const* summary_hosts, int summary_count, void* workspace,
                           size_t workspace_bytes, spvd_plan_graph_t** out);
int spvd_plan_graph_launch(spvd_plan_graph_t* g, const float* pc, void* const* events, spvd_stream_t stream,
                           spvd_stream_t side_stream);
void spvd_plan_graph_destroy(spvd_plan_graph_t* g);

/* ------------------------------------------------------------------------------------------------
 * M2  native inference executor: SPVUnet.forward (models/spvd.py:432-457, eval mode) as one call.
 * The host compiles the module tree once into spvd_op_t records whose operands are byte offsets
 * into one activation arena; row counts are resolved per call from the planned geometry. */
#define SPVD_OP_SCATTER_MEAN 0 /* src0 [points,c0] -> dst [level rows,c0]; aux = counts scratch   */
#define SPVD_OP_DEVOX 1        /* src0 [level rows,c0] -> dst [points,c0]                          */
#define SPVD_OP_CONV 2         /* src0 [rows_in,c0] -> dst [rows_out,c1] (+bias, +src1 residual)   */
#define SPVD_OP_AFFINE 3       /* dst = act(src0 * p0[c] + p1[c])                                  */
#define SPVD_OP_FILM_AFFINE 4  /* src1 = tt table (row stride i0 floats, column offset i1)         */
#define SPVD_OP_CAT 5          /* dst = [src0 (c0) | src1 (c1)]                                    */
#define SPVD_OP_LAYERNORM 6    /* p0 = gamma, p1 = beta                                            */
#define SPVD_OP_ATTENTION 7    /* src0 = qkv [rows,3*c1] -> dst [rows,c1]; i0 = heads              */
#define SPVD_OP_TIME_EMBED 8   /* dst [batch,c1]; c0 = n_temb; w = params, p0 = class table or NULL,
                                  p1 = t (int64 [batch]), bias = class ids (int64 [batch]) or NULL  */
#define SPVD_OP_CAT_AFFINE 9   /* dst = f16 act([src0|src1]*p0+p1), aux = f16 [src0|src1]          */
#define SPVD_OP_CONV_SP 10     /* spvd_conv_sp_f16: src0 f16 [rows_in,c0] (+ src2 f16 [rows_out,c2]) -> v = conv + bias
                                  + src1 residual, FiLM(src3); dst = v as fp32 (or -1), aux = f16 act(v * p0 + p1) (or
                                  -1; p0 NULL = no affine, SPVD_FLAG_SILU = SiLU)                                  */

#define SPVD_MAP_NONE 0 /* kernel size 1 / nn.Linear */
#define SPVD_MAP_K3 1
#define SPVD_MAP_DOWN 2 /* level -> level+1 */
#define SPVD_MAP_UP 3   /* level+1 -> level */

#define SPVD_ROWS_POINTS (-1)
#define SPVD_ROWS_BATCH (-2)

#define SPVD_FLAG_SILU 1   /* AFFINE / FILM_AFFINE: apply SiLU                                     */
#define SPVD_FLAG_ROUND 2  /* producer rounds its output to TF32                                    */
#define SPVD_FLAG_A_TF32 4 /* CONV: src0 is TF32-representable (cp.async gather)                   */
#define SPVD_FLAG_SIMT 8   /* CONV: force the fp32 SIMT kernel                                     */
#define SPVD_FLAG_F16 16   /* CONV: f16 operands (w_packed = f16 image; with A_TF32: src0 is f16);
                              AFFINE / FILM_AFFINE / LAYERNORM / ATTENTION: write dst as f16       */

typedef struct spvd_level_geom {
  const int32_t* coords; /* [n,4] */
  const int32_t* kmap3;
  const uint32_t* mask3;
  const int32_t* kmap_down;
  const uint32_t* mask_down;
  const int32_t* kmap_up;
  const uint32_t* mask_up;
  const int32_t* p2v;       /* [n_points] or NULL */
  const int32_t* devox_idx; /* [n_points,8] or NULL */
  const float* devox_w;
  const int32_t* batch_offsets; /* [n_batch+1] or NULL */
  const int32_t* pair_in;       /* pair-major view of kmap3 or NULL */
  const int32_t* pair_out;
  const int32_t* pair_tile_k;
  const int32_t* dpair_in;      /* pair-major view of kmap_down (fine rows) or NULL */
  const int32_t* dpair_out;     /* coarse rows */
  const int32_t* dpair_tile_k;
  const float* p2v_w;       /* per-point mean weights (spvd_point_mean_weights) or NULL */
  const int32_t* p2v_cnt;   /* points grouped by voxel (spvd_group_points_by_voxel) or NULL */
  const int32_t* p2v_start;
  const int32_t* p2v_order;
  const int32_t* p2v_units;
  const int32_t* kmap3_sorted;   /* mask-sorted views (spvd_kmap_sort) or NULL */
  const uint32_t* mask3_sorted;
  const int32_t* perm3;
  const int32_t* kmap_down_sorted;
  const uint32_t* mask_down_sorted;
  const int32_t* perm_down;
  const int32_t* kmap_up_sorted;
  const uint32_t* mask_up_sorted;
  const int32_t* perm_up;
  int32_t n, ld3, ld_down, ld_up, max_per_batch, n_pair_tiles; /* n_pair_tiles > 0: 3^3 convs run pair-major */
  int32_t n_dpair_tiles, pad;   /* > 0: the stride-2 / transposed convs of this level run pair-major */
} spvd_level_geom_t;

typedef struct spvd_op {
  int32_t op, flags, level, map_kind;
  int32_t rows_in, rows_out; /* level index, SPVD_ROWS_POINTS or SPVD_ROWS_BATCH */
  int32_t c0, c1, kvol, i0, i1, pad;
  int64_t src0, src1, dst, aux; /* byte offsets into the arena, -1 = none */
  const float* w;
  const float* w_packed;
  const float* bias;
  const float* p0;
  const float* p1;
  /* SPVD_OP_CONV_SP only: src2 = second operand source (f16 [rows_out, c2], identity rows; -1 = none) with its packed
   * k = 1 weight image w2_packed; src3 = FiLM table (row stride i0 floats, column offset i1; -1 = none) */
  int64_t src2, src3;
  const float* w2_packed;
  int32_t c2, pad2;
} spvd_op_t;

/* conv_events: NULL, or 2 CUDA events (cudaEvent_t) per SPVD_OP_CONV in program order, recorded
 * around each convolution launch (bench.py's live roofline measurement). */
int spvd_program_run(const spvd_op_t* ops, int n_ops, const spvd_level_geom_t* levels, int n_levels,
                     int n_points, int n_batch, void* arena, size_t arena_bytes,
                     void* const* conv_events, spvd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * TR  training-mode glue (csrc/train.cu): what a training step of the denoiser needs between the sparse convolutions,
 * so that forward + backward run on this library only (train_generation.py:98-112 drives it; the modules are
 * models/ddpm_unet_attn.py:24-65, models/modules/graph.py:19-41, models/modules/transformer.py:40-67,102-125).
 * `workspace` everywhere: >= 2 * c doubles (1 for spvd_mse_loss), ZERO on entry; every call leaves it zero again. */
/* nn.BatchNorm1d in training mode (+ SiLU when act & 1): batch statistics over `rows`, y = act(x * scale + shift);
 * save_mean / save_invstd / scale / shift [c] are outputs; running_mean / running_var (may be NULL) are updated in place
 * with `momentum` and the unbiased variance, exactly like torch. */
int spvd_bn_train_fwd(const float* x, long long rows, int c, const float* gamma, const float* beta, float eps,
                      float momentum, float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                      float* scale, float* shift, int act, float* y, double* workspace, spvd_stream_t stream);
int spvd_bn_train_bwd(const float* x, const float* gy, long long rows, int c, const float* gamma, const float* beta,
                      const float* save_mean, const float* save_invstd, int act, float* gx, float* ggamma, float* gbeta,
                      double* workspace, spvd_stream_t stream);
/* out[c] = sum over rows of g[., c]  (bias gradients) */
int spvd_colsum(const float* g, long long rows, int c, float* out, double* workspace, spvd_stream_t stream);
/* backward of y = (1 + tt[b, c]) * x + tt[b, C + c] (spvd_film_fwd): gx, and gtt [n_batch, ldg] += the per-cloud sums
 * (gtt zero-filled by the caller); rows of cloud b are [batch_offsets[b], batch_offsets[b+1]). */
int spvd_film_bwd(const float* x, const float* gy, const float* tt, int ldt, const int32_t* batch_offsets, int n_batch,
                  int max_rows_per_cloud, int c, float* gx, float* gtt, int ldg, spvd_stream_t stream);
int spvd_layernorm_train_fwd(const float* x, const float* gamma, const float* beta, int rows, int c, float eps, float* y,
                             float* mean, float* rstd, spvd_stream_t stream);
int spvd_layernorm_bwd(const float* x, const float* gy, const float* gamma, const float* mean, const float* rstd, int rows,
                       int c, float* gx, float* ggamma, float* gbeta, double* workspace, spvd_stream_t stream);
/* spvd_attention_fwd that also saves the log-sum-exp [rows, heads]; backward: gqkv [rows, 3c], delta [rows, heads] scratch */
int spvd_attention_train_fwd(const float* qkv, const int32_t* batch_offsets, int n_batch, int max_len, int c, int heads,
                             float* out, float* lse, spvd_stream_t stream);
int spvd_attention_bwd(const float* qkv, const float* out, const float* gout, const float* lse, const int32_t* batch_offsets,
                       int n_batch, int max_len, int c, int heads, float* gqkv, float* delta, spvd_stream_t stream);
/* The whole backward of one sparse convolution (or nn.Linear = kernel size 1, map NULL) from one call: gx = data gradient
 * (the forward kernels on the transposed weights and the gradient map gmap_t: same map with flipped offsets for a
 * submanifold conv (flip = 1), the transposed map for strided convs), gw = weight gradient (tcgen05 over the pair lists
 * pair_in / pair_out / pair_tile_k of spvd_kmap_pairs -- NULL with kvol == 1: identity pairs -- or the fp32 SIMT kernel on
 * map_t when the widths are not multiples of 32 / allow_tensor_cores == 0), gbias = column sums of g.  Any of gx / gw /
 * gbias may be NULL.  x_is_tf32: the forward input is already TF32-representable; gw_transposed: write gw as
 * [kvol, cout, cin] (the layout of nn.Linear.weight).  workspace >= spvd_conv_bwd_workspace_bytes(...);
 * red_workspace = the zeroed fp64 scratch of this section (>= 2 * cout doubles). */
typedef struct spvd_conv_bwd_args {
  const float* x;        /* [n_in, cin] forward input */
  const float* g;        /* [n_out, cout] gradient of the output */
  const float* w;        /* [kvol, cin, cout] forward weights */
  const int32_t* map_t;  /* forward map [kvol, map_ld] (SIMT weight gradient) or NULL */
  const int32_t* gmap_t; /* gradient map [kvol, gmap_ld] or NULL (kernel size 1) */
  const uint32_t* gmap_mask;
  const int32_t* pair_in;
  const int32_t* pair_out;
  const int32_t* pair_tile_k;
  float* gx;
  float* gw;
  float* gbias;
  void* workspace;
  double* red_workspace;
  const float* w_packed_t; /* optional: transposed tile image from spvd_conv_pack_weights_train (else built per call) */
  size_t workspace_bytes;
  int32_t map_ld, gmap_ld, n_pair_tiles, n_in, n_out, kvol, cin, cout, flip, x_is_tf32, gw_transposed, allow_tensor_cores;
  int32_t n_pairs;         /* valid pairs in the lists (0 = unknown): the data gradient runs pair-major when n_pairs < 10 n_in */
  int32_t w_packed_t_flip; /* the flip_t w_packed_t was built with; it is used only if that is what this call needs */
} spvd_conv_bwd_args_t;
size_t spvd_conv_bwd_workspace_bytes(int n_in, int n_out, int kvol, int cin, int cout);
/* Training: every per-step derivative of one weight tensor in one launch -- the forward tile image wp_out (layout of
 * spvd_conv_pack_weights), the transposed image wtp_out for the data gradient (offsets reversed when flip_t: the dense
 * submanifold data gradient; 0 for the pair-major one and for strided maps) and, when w is an nn.Linear weight [cout, cin]
 * (w_is_linear), its conv layout w3_out [1, cin, cout].  Any output may be NULL. */
int spvd_conv_pack_weights_train(const float* w, int kvol, int cin, int cout, int w_is_linear, int flip_t, float* w3_out,
                                 float* wp_out, float* wtp_out, spvd_stream_t stream);
int spvd_conv_bwd(const spvd_conv_bwd_args_t* args, spvd_stream_t stream);
/* torchsparse.cat (models/spvd.py:232) for any channel counts, and its backward (channel split) */
int spvd_cat_any_fwd(const float* a, const float* b, long long rows, int ca, int cb, float* y, spvd_stream_t stream);
int spvd_split_fwd(const float* g, long long rows, int ca, int cb, float* ga, float* gb, spvd_stream_t stream);
int spvd_silu_fwd(const float* x, long long n, float* y, spvd_stream_t stream);
int spvd_silu_bwd(const float* x, const float* gy, long long n, float* gx, spvd_stream_t stream);
int spvd_add_fwd(const float* a, const float* b, long long n, float* y, spvd_stream_t stream);
/* out[b] = table[idx[b]] (+ base[b]): the class embedding added to the time embedding (ConditionalGeneration.ipynb cell 4);
 * backward: gtable[idx[b]] += g[b] (gtable zero-filled by the caller) */
int spvd_embedding_fwd(const float* table, const int64_t* idx, const float* base, int n_batch, int e, float* out,
                       spvd_stream_t stream);
int spvd_embedding_bwd(const float* g, const int64_t* idx, int n_batch, int e, float* gtable, spvd_stream_t stream);
/* models/utils.py:15-19: out[b] = [sin(t_b * f_i), cos(t_b * f_i)], f_i = exp(-ln(1e4) * i / (d/2 - 1)) */
int spvd_timestep_embedding(const int64_t* t, int n_batch, int d, float* out, spvd_stream_t stream);
/* nn.MSELoss (train_generation.py:104): loss[0] = mean((pred - target)^2); gpred (may be NULL) = grad_scale * d loss / d pred */
int spvd_mse_loss(const float* pred, const float* target, long long n, float grad_scale, float* loss, float* gpred,
                  double* workspace, spvd_stream_t stream);

/* Optimizer tail of train_generation.py:106-112 over flat fp32 buffers of n elements (n % 4 == 0, 16-byte aligned; the
 * gradient buffer of spvd_b200.parallel.GradientAllReduce and the parameter / moment buffers of spvd_b200.optim.FlatAdam):
 * spvd_sumsq: *acc = sum x^2 (fp64; zeroed by the call).
 * spvd_adam_step: torch.optim.Adam (utils/... LearnerCallbacks use Adam; no amsgrad) at step `step` (1-based) on
 * g' = g * scale, scale = pending_scale * min(1, max_norm / (pending_scale * sqrt(*sumsq) + 1e-6)) -- clip_grad_norm_
 * (utils/callbacks.py:62-63) and the data-parallel 1/world average folded into the same pass; sumsq NULL or
 * max_norm <= 0 = no clipping.  norm_out (may be NULL) receives the norm of the averaged gradient. */
int spvd_sumsq(const float* x, long long n, double* acc, spvd_stream_t stream);
int spvd_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, const double* sumsq, float max_norm, float pending_scale, float* norm_out,
                   spvd_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * N4  evaluation distances between point clouds xyz1 [B,n,3], xyz2 [B,m,3] (fp32, row-major).
 * Chamfer (metrics/chamfer_dist/chamfer.cu:15-229): dist1[b,i] = min_j |xyz1_i - xyz2_j|^2 with its first argmin idx1, and
 * the same from xyz2 to xyz1; the backward scatters 2 g (a - b) to both ends of every nearest-neighbour pair.
 * Approximate earth mover's distance (metrics/PyTorchEMD/cuda/emd_kernel.cu:26-330): approxmatch = auction-style soft
 * matching over the levels -4^7 .. -4^-1, 0 into match [B,m,n]; matchcost = sum match * squared distance, and its
 * gradients w.r.t. both clouds.  workspace: spvd_emd_workspace_floats(B, n, m) floats. */
int spvd_chamfer_fwd(const float* xyz1, const float* xyz2, int n_batch, int n, int m, float* dist1, float* dist2,
                     int32_t* idx1, int32_t* idx2, spvd_stream_t stream);
int spvd_chamfer_bwd(const float* xyz1, const float* xyz2, const int32_t* idx1, const int32_t* idx2, const float* gdist1,
                     const float* gdist2, int n_batch, int n, int m, float* gxyz1, float* gxyz2, spvd_stream_t stream);
size_t spvd_emd_workspace_floats(int n_batch, int n, int m);
int spvd_emd_approxmatch(const float* xyz1, const float* xyz2, int n_batch, int n, int m, float* match, float* workspace,
                         spvd_stream_t stream);
int spvd_emd_matchcost_fwd(const float* xyz1, const float* xyz2, const float* match, int n_batch, int n, int m, float* cost,
                           spvd_stream_t stream);
int spvd_emd_matchcost_bwd(const float* grad_cost, const float* xyz1, const float* xyz2, const float* match, int n_batch, int n,
                           int m, float* grad1, float* grad2, spvd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPVD_B200_H */
