// Geometry planner: every coordinate-only structure of one forward (voxel pyramid, hash tables,
// kernel maps, point<->voxel indices, trilinear weights, per-cloud counts) from ONE C call into
// caller-owned buffers, ending with the single device->host read of the level sizes.  Replaces the
// ~20 synchronising steps of the reference (torch.unique / scatter_mean sizes / kernel-map builds,
// SURVEY.md section 3.1) with ~150 back-to-back launches and one 1 KB copy.
#include "common.cuh"

namespace spvd {

__global__ void k_batch_offsets(const int* __restrict__ counts, int nb, int* __restrict__ offsets) {
  // nb is small (clouds per batch): one thread
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int run = 0;
    offsets[0] = 0;
    for (int b = 0; b < nb; ++b) {
      run += counts[b];
      offsets[b + 1] = run;
    }
  }
}

__global__ void k_clamp_min0(int* __restrict__ idx, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && idx[i] < 0) idx[i] = 0;
}

}  // namespace spvd

using namespace spvd;

extern "C" {
int spvd_point_voxel_coords(const float*, int, float, float, int, float*, int32_t*, spvd_stream_t);
size_t spvd_unique_workspace_bytes(int);
int spvd_unique_coords(const int32_t*, int, const int32_t*, int32_t*, int32_t*, int32_t*, void*, size_t, spvd_stream_t);
int spvd_downsample_coords(const int32_t*, int, const int32_t*, int, int32_t*, int32_t*, int32_t*, void*, size_t,
                           spvd_stream_t);
int spvd_hash_build(const int32_t*, int, const int32_t*, int, int64_t*, int32_t*, int, int32_t*, spvd_stream_t);
int spvd_kmap_subm3(const int32_t*, int, const int32_t*, const int64_t*, const int32_t*, int, int32_t*, int, uint32_t*,
                    spvd_stream_t);
int spvd_kmap_down2(const int32_t*, int, const int32_t*, const int64_t*, const int32_t*, int, int32_t*, int, uint32_t*,
                    int32_t*, int, uint32_t*, int32_t*, int32_t*, spvd_stream_t);
int spvd_point_voxel_query(const float*, int, int, const int64_t*, const int32_t*, int, int32_t*, spvd_stream_t);
int spvd_devox_query(const float*, int, int, const int64_t*, const int32_t*, int, int32_t*, float*, spvd_stream_t);
int spvd_batch_counts(const int32_t*, int, const int32_t*, int, int32_t*, spvd_stream_t);
int spvd_kmap_pairs(const int32_t*, int, int, const int32_t*, int, int, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*,
                    int32_t*, spvd_stream_t);
int spvd_point_mean_weights(const int32_t*, int, int, int32_t*, float*, spvd_stream_t);
int spvd_group_points_by_voxel(const int32_t*, int, int, const int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*,
                               spvd_stream_t);
int spvd_segmented_unique(const int32_t*, int, const int32_t*, int, int, int, int, int32_t*, int32_t*, int32_t*, int32_t*,
                          int32_t*, void*, size_t, spvd_stream_t);
int spvd_kmap_sort(const int32_t*, int, int, const int32_t*, int, int32_t*, uint32_t*, int32_t*, void*, size_t, spvd_stream_t);
}

#define SPVD_TRY(call)            \
  do {                            \
    int rc_ = (call);             \
    if (rc_ != SPVD_OK) return rc_; \
  } while (0)

// A second stream + two events per (host thread, device): inside one level group the point queries (point->voxel index,
// trilinear corners) depend only on the hash table and run beside the 3^3 map and its pair list.  Works the same when the
// caller's stream is being captured (the fork / join become graph edges).
struct ForkCtx {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static int fork_ctx(ForkCtx** out) {
  static thread_local ForkCtx ctx[64];
  int dev = 0;
  SPVD_CUDA(cudaGetDevice(&dev));
  SPVD_CHECK_ARG(dev >= 0 && dev < 64, "device index %d out of range", dev);
  ForkCtx& c = ctx[dev];
  if (!c.stream) {
    SPVD_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    SPVD_CUDA(cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming));
    SPVD_CUDA(cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming));
  }
  *out = &c;
  return SPVD_OK;
}

// phases: 1 = the coordinate chain (voxel sets, hash tables, stride-2 maps), 2 = what hangs off it per level (3^3 maps,
// pair lists, point queries) and the summary copy; the coarser levels need only phase 1 of the finer ones
static int plan_range(const float* pc, int n_points, int n_batch, int max_cloud_rows, float init_res, float after_res,
                      int scale_mode, int downsample_mode, const spvd_plan_level_t* levels, int n_levels, int level_begin,
                      int level_end, int phases, int ld, float* fcoords, int32_t* icoords, int32_t* status,
                      const int32_t* summary_dev, int32_t* summary_host, int summary_count, int sync, void* workspace,
                      size_t workspace_bytes, spvd_stream_t stream) {
  SPVD_CHECK_ARG(pc && levels && fcoords && icoords && status && workspace && n_points > 0 && n_levels > 0,
                 "spvd_plan_build: bad arguments");
  SPVD_CHECK_ARG(0 <= level_begin && level_begin < level_end && level_end <= n_levels, "spvd_plan_build: bad level range [%d,%d)",
                 level_begin, level_end);
  SPVD_CHECK_ARG(ld % 128 == 0 && ld >= n_points, "spvd_plan_build: ld %d must be a multiple of 128 >= n_points", ld);
  cudaStream_t s = (cudaStream_t)stream;
  if (level_begin == 0 && (phases & 1)) {
    SPVD_CUDA(cudaMemsetAsync(status, 0, 4, s));
    SPVD_TRY(spvd_point_voxel_coords(pc, n_points, init_res, after_res, scale_mode, fcoords, icoords, stream));
  }
  for (int i = level_begin; i < level_end && (phases & 1); ++i) {
    const spvd_plan_level_t& L = levels[i];
    SPVD_CHECK_ARG(L.coords && L.count && L.keys && L.vals && L.capacity >= 2 * n_points,
                   "spvd_plan_build: level %d lacks coords / count / hash buffers", i);
    const bool seg = max_cloud_rows > 0 && max_cloud_rows <= 16384 && n_batch > 0 && L.batch_counts;
    if (seg) {
      // one CTA per cloud sorts in shared memory; per-cloud counts / offsets of the level come for free
      SPVD_TRY(spvd_segmented_unique(i == 0 ? icoords : levels[i - 1].coords, n_points, i == 0 ? nullptr : levels[i - 1].count,
                                     n_batch, max_cloud_rows, i > 0, downsample_mode, L.coords, L.count, L.batch_counts,
                                     L.batch_offsets, status, workspace, workspace_bytes, stream));
    } else if (i == 0) {
      SPVD_TRY(spvd_unique_coords(icoords, n_points, nullptr, L.coords, L.count, status, workspace, workspace_bytes, stream));
    } else {
      SPVD_TRY(spvd_downsample_coords(levels[i - 1].coords, n_points, levels[i - 1].count, downsample_mode, L.coords,
                                      L.count, status, workspace, workspace_bytes, stream));
    }
    SPVD_TRY(spvd_hash_build(L.coords, n_points, L.count, SPVD_ORDER_BXYZ, L.keys, L.vals, L.capacity, status, stream));
    if (i > 0) {
      const spvd_plan_level_t& F = levels[i - 1];
      if (F.kmap_down || F.kmap_up)
        SPVD_TRY(spvd_kmap_down2(F.coords, n_points, F.count, L.keys, L.vals, L.capacity, F.kmap_down, ld, F.mask_down,
                                 F.kmap_up, ld, F.mask_up, nullptr, nullptr, stream));
      // (scratch of the COARSE level: the fine level's own scratch may be in use by its 3^3 pair list on another stream)
      if (F.kmap_down && F.dpair_in && F.dpair_out && F.dpair_tile_k && F.dpair_counts && L.pair_scratch)
        SPVD_TRY(spvd_kmap_pairs(F.kmap_down, ld, n_points, L.count, 8, F.dpair_tile_cap, F.dpair_in, F.dpair_out,
                                 F.dpair_tile_k, F.dpair_counts, F.dpair_counts + 1, L.pair_scratch, stream));
    }
  }
  if (phases & 2) {
    ForkCtx* fk = nullptr;
    SPVD_TRY(fork_ctx(&fk));
    bool forked = false;
    for (int i = level_begin; i < level_end; ++i) forked = forked || levels[i].p2v || (levels[i].devox_idx && levels[i].devox_w);
    if (forked) {
      // point queries of the group on the side branch
      SPVD_CUDA(cudaEventRecord(fk->fork, s));
      SPVD_CUDA(cudaStreamWaitEvent(fk->stream, fk->fork, 0));
      for (int i = level_begin; i < level_end; ++i) {
        const spvd_plan_level_t& L = levels[i];
        const int stride = 1 << i;
        if (L.p2v) {
          SPVD_TRY(spvd_point_voxel_query(fcoords, n_points, stride, L.keys, L.vals, L.capacity, L.p2v, (spvd_stream_t)fk->stream));
          // misses go to row 0: idx_query.clamp_(0), models/sparse_utils.py:93
          k_clamp_min0<<<cdiv(n_points, 256), 256, 0, fk->stream>>>(L.p2v, n_points);
          if (L.p2v_cnt && L.p2v_start && L.p2v_cursor && L.p2v_order && L.p2v_units)
            SPVD_TRY(spvd_group_points_by_voxel(L.p2v, n_points, n_points, L.count, L.p2v_cnt, L.p2v_start, L.p2v_cursor,
                                                L.p2v_order, L.p2v_units, (spvd_stream_t)fk->stream));
          else if (L.p2v_w && L.p2v_cnt)
            SPVD_TRY(spvd_point_mean_weights(L.p2v, n_points, n_points, L.p2v_cnt, L.p2v_w, (spvd_stream_t)fk->stream));
        }
        if (L.devox_idx && L.devox_w)
          SPVD_TRY(spvd_devox_query(fcoords, n_points, stride, L.keys, L.vals, L.capacity, L.devox_idx, L.devox_w,
                                    (spvd_stream_t)fk->stream));
      }
      SPVD_CUDA(cudaEventRecord(fk->join, fk->stream));
    }
    for (int i = level_begin; i < level_end; ++i) {
      const spvd_plan_level_t& L = levels[i];
      if (L.kmap3) SPVD_TRY(spvd_kmap_subm3(L.coords, n_points, L.count, L.keys, L.vals, L.capacity, L.kmap3, ld, L.mask3, stream));
      if (L.kmap3 && L.kmap3_sorted && L.mask3_sorted && L.perm3 && L.sort_ws)   // mask-sorted view for the persistent conv
        SPVD_TRY(spvd_kmap_sort(L.kmap3, ld, n_points, L.count, 27, L.kmap3_sorted, L.mask3_sorted, L.perm3, L.sort_ws,
                                (size_t)L.sort_ws_bytes, stream));
      if (L.kmap3 && L.pair_in && L.pair_out && L.pair_tile_k && L.pair_counts && L.pair_scratch)
        SPVD_TRY(spvd_kmap_pairs(L.kmap3, ld, n_points, L.count, 27, L.pair_tile_cap, L.pair_in, L.pair_out, L.pair_tile_k,
                                 L.pair_counts, L.pair_counts + 1, L.pair_scratch, stream));
      const bool seg = max_cloud_rows > 0 && max_cloud_rows <= 16384 && n_batch > 0 && L.batch_counts;
      if (L.batch_counts && !seg) {
        SPVD_TRY(spvd_batch_counts(L.coords, n_points, L.count, n_batch, L.batch_counts, stream));
        if (L.batch_offsets) k_batch_offsets<<<1, 32, 0, s>>>(L.batch_counts, n_batch, L.batch_offsets);
      }
    }
    if (forked) SPVD_CUDA(cudaStreamWaitEvent(s, fk->join, 0));
  }
  SPVD_LAUNCH_CHECK();
  if ((phases & 2) && summary_dev && summary_host && summary_count > 0) {
    SPVD_CUDA(cudaMemcpyAsync(summary_host, summary_dev, (size_t)summary_count * 4, cudaMemcpyDeviceToHost, s));
    if (sync) SPVD_CUDA(cudaStreamSynchronize(s));
  }
  return SPVD_OK;
}

extern "C" int spvd_plan_build_range(const float* pc, int n_points, int n_batch, int max_cloud_rows, float init_res,
                                     float after_res, int scale_mode, int downsample_mode, const spvd_plan_level_t* levels,
                                     int n_levels, int level_begin, int level_end, int ld, float* fcoords, int32_t* icoords,
                                     int32_t* status, const int32_t* summary_dev, int32_t* summary_host, int summary_count,
                                     int sync, void* workspace, size_t workspace_bytes, spvd_stream_t stream) {
  return plan_range(pc, n_points, n_batch, max_cloud_rows, init_res, after_res, scale_mode, downsample_mode, levels, n_levels,
                    level_begin, level_end, 3, ld, fcoords, icoords, status, summary_dev, summary_host, summary_count, sync,
                    workspace, workspace_bytes, stream);
}

extern "C" int spvd_plan_build(const float* pc, int n_points, int n_batch, int max_cloud_rows, float init_res, float after_res,
                               int scale_mode, int downsample_mode, const spvd_plan_level_t* levels, int n_levels,
                               int ld, float* fcoords, int32_t* icoords, int32_t* status, const int32_t* summary_dev,
                               int32_t* summary_host, int summary_count, void* workspace, size_t workspace_bytes,
                               spvd_stream_t stream) {
  return spvd_plan_build_range(pc, n_points, n_batch, max_cloud_rows, init_res, after_res, scale_mode, downsample_mode, levels,
                               n_levels, 0, n_levels, ld, fcoords, icoords, status, summary_dev, summary_host, summary_count, 1,
                               workspace, workspace_bytes, stream);
}

// Level groups [bounds[g], bounds[g+1]) planned with ONE host call and no host wait: group 0 on `stream`, the others on
// `side_stream` behind group 0; events[g] (cudaEvent_t) is recorded after group g's summary copy into summary_hosts[g].
extern "C" int spvd_plan_build_groups(const float* pc, int n_points, int n_batch, int max_cloud_rows, float init_res,
                                      float after_res, int scale_mode, int downsample_mode, const spvd_plan_level_t* levels,
                                      int n_levels, const int32_t* bounds, int n_groups, int ld, float* fcoords,
                                      int32_t* icoords, int32_t* status, const int32_t* summary_dev,
                                      int32_t* const* summary_hosts, int summary_count, void* const* events, void* workspace,
                                      size_t workspace_bytes, spvd_stream_t stream, spvd_stream_t side_stream) {
  SPVD_CHECK_ARG(bounds && summary_hosts && events && n_groups >= 1 && bounds[0] == 0 && bounds[n_groups] == n_levels,
                 "spvd_plan_build_groups: bounds must run from 0 to n_levels");
  for (int g = 0; g < n_groups; ++g) {
    SPVD_CHECK_ARG(summary_hosts[g] && events[g], "spvd_plan_build_groups: group %d lacks a summary buffer / event", g);
    spvd_stream_t st = g == 0 ? stream : side_stream;
    SPVD_TRY(spvd_plan_build_range(pc, n_points, n_batch, max_cloud_rows, init_res, after_res, scale_mode, downsample_mode,
                                   levels, n_levels, bounds[g], bounds[g + 1], ld, fcoords, icoords, status, summary_dev,
                                   summary_hosts[g], summary_count, 0, workspace, workspace_bytes, st));
    SPVD_CUDA(cudaEventRecord((cudaEvent_t)events[g], (cudaStream_t)st));
    if (g == 0 && n_groups > 1) SPVD_CUDA(cudaStreamWaitEvent((cudaStream_t)side_stream, (cudaEvent_t)events[0], 0));
  }
  return SPVD_OK;
}

// ------------------------------------------------------------------------------------------------
// The planner as CUDA graphs: its launch sequence depends only on capacities (every grid is sized by n_points, the row
// counts stay on the device), so one capture per workspace replaces ~65 launches by one graph launch per level group.
// ------------------------------------------------------------------------------------------------
struct spvd_plan_graph {
  int n_groups = 0;
  size_t pc_bytes = 0;
  float* pc_stage = nullptr;
  cudaGraphExec_t exec[8] = {};
  cudaGraphExec_t exec0_tail = nullptr;   // with several groups: exec[0] = coordinate chain of group 0, this = the rest of it
  cudaEvent_t chain_done = nullptr;       // the coarser groups start behind the coordinate chain, not behind all of group 0
  cudaStream_t capture = nullptr;
};

extern "C" void spvd_plan_graph_destroy(spvd_plan_graph* g) {
  if (!g) return;
  for (int i = 0; i < g->n_groups; ++i)
    if (g->exec[i]) cudaGraphExecDestroy(g->exec[i]);
  if (g->exec0_tail) cudaGraphExecDestroy(g->exec0_tail);
  if (g->chain_done) cudaEventDestroy(g->chain_done);
  if (g->capture) cudaStreamDestroy(g->capture);
  delete g;
}

extern "C" int spvd_plan_graph_create(int n_points, int n_batch, int max_cloud_rows, float init_res, float after_res,
                                      int scale_mode, int downsample_mode, const spvd_plan_level_t* levels, int n_levels,
                                      const int32_t* bounds, int n_groups, int ld, float* pc_stage, float* fcoords,
                                      int32_t* icoords, int32_t* status, const int32_t* summary_dev,
                                      int32_t* const* summary_hosts, int summary_count, void* workspace,
                                      size_t workspace_bytes, spvd_plan_graph** out) {
  SPVD_CHECK_ARG(out && pc_stage && bounds && summary_hosts && n_groups >= 1 && n_groups <= 8 && bounds[0] == 0 &&
                     bounds[n_groups] == n_levels,
                 "spvd_plan_graph_create: bad arguments (1..8 groups, bounds from 0 to n_levels)");
  *out = nullptr;
  spvd_plan_graph* g = new spvd_plan_graph();
  g->pc_stage = pc_stage;
  g->pc_bytes = (size_t)n_points * 4 * sizeof(float);
  cudaError_t e = cudaStreamCreateWithFlags(&g->capture, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("spvd_plan_graph_create: cudaStreamCreate failed: %s", cudaGetErrorString(e));
    delete g;
    return SPVD_ERR_CUDA;
  }
  if (n_groups > 1) e = cudaEventCreateWithFlags(&g->chain_done, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    set_error("spvd_plan_graph_create: cudaEventCreate failed: %s", cudaGetErrorString(e));
    spvd_plan_graph_destroy(g);
    return SPVD_ERR_CUDA;
  }
  // graphs: one per level group; group 0 of a split plan is two (coordinate chain | the rest)
  for (int i = 0; i < n_groups + (n_groups > 1 ? 1 : 0); ++i) {
    const int grp = i < n_groups ? i : 0;
    const int phases = n_groups == 1 ? 3 : (i == 0 ? 1 : (i == n_groups ? 2 : 3));
    e = cudaStreamBeginCapture(g->capture, cudaStreamCaptureModeRelaxed);
    int rc = SPVD_OK;
    if (e == cudaSuccess)
      rc = plan_range(pc_stage, n_points, n_batch, max_cloud_rows, init_res, after_res, scale_mode, downsample_mode, levels,
                      n_levels, bounds[grp], bounds[grp + 1], phases, ld, fcoords, icoords, status, summary_dev,
                      summary_hosts[grp], summary_count, 0, workspace, workspace_bytes, (spvd_stream_t)g->capture);
    cudaGraph_t graph = nullptr;
    cudaError_t e2 = e == cudaSuccess ? cudaStreamEndCapture(g->capture, &graph) : e;
    cudaGraphExec_t* slot = i < n_groups ? &g->exec[i] : &g->exec0_tail;
    if (rc == SPVD_OK && e2 == cudaSuccess) e2 = cudaGraphInstantiate(slot, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (i < n_groups) g->n_groups = i + 1;
    if (rc != SPVD_OK || e2 != cudaSuccess) {
      if (rc == SPVD_OK) set_error("spvd_plan_graph_create: capture of level group %d failed: %s", grp, cudaGetErrorString(e2));
      cudaGetLastError();
      spvd_plan_graph_destroy(g);
      return rc != SPVD_OK ? rc : SPVD_ERR_CUDA;
    }
  }
  *out = g;
  return SPVD_OK;
}

// group 0 on `stream`, the others on `side_stream` behind it; events[i] (cudaEvent_t) is recorded after group i
extern "C" int spvd_plan_graph_launch(spvd_plan_graph* g, const float* pc, void* const* events, spvd_stream_t stream,
                                      spvd_stream_t side_stream) {
  SPVD_CHECK_ARG(g && pc && events, "spvd_plan_graph_launch: bad arguments");
  cudaStream_t s = (cudaStream_t)stream, side = (cudaStream_t)side_stream;
  SPVD_CUDA(cudaMemcpyAsync(g->pc_stage, pc, g->pc_bytes, cudaMemcpyDeviceToDevice, s));
  for (int i = 0; i < g->n_groups; ++i) {
    SPVD_CHECK_ARG(events[i], "spvd_plan_graph_launch: group %d has no event", i);
    cudaStream_t st = i == 0 ? s : side;
    SPVD_CUDA(cudaGraphLaunch(g->exec[i], st));
    if (i == 0 && g->n_groups > 1) {
      // the coarser groups wait for group 0's coordinate chain only and run beside its maps and point queries
      SPVD_CUDA(cudaEventRecord(g->chain_done, s));
      SPVD_CUDA(cudaStreamWaitEvent(side, g->chain_done, 0));
      SPVD_CUDA(cudaGraphLaunch(g->exec0_tail, s));
    }
    SPVD_CUDA(cudaEventRecord((cudaEvent_t)events[i], st));
  }
  return SPVD_OK;
}
