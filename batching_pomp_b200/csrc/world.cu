// world.cu — device-resident collision worlds, batched isValid and the vertex prune predicate.
// Reference citations are relative to /root/reference/.
#include <cfloat>
#include <cmath>
#include <cstring>

#include "world.cuh"
#include "fastpath.cuh"
#include "fastpath_host.hpp"

// ---------------------------------------------------------------------------
// World upload
// ---------------------------------------------------------------------------
extern "C" int bpomp_world_set_image(bpomp_ctx* ctx, const uint8_t* pix, int width, int height, int threshold) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (ctx->dim != 2) return bp_fail(ctx, BPOMP_ERR_INVALID, "the image world needs dim == 2 (ctx has %d)", ctx->dim);
  if (!pix || width <= 0 || height <= 0) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad image");
  // Pack to 8x8-pixel tiles of one bit per pixel ("free" = pix >= threshold): a 4096^2 map is
  // 2 MB and stays L2-resident; one 8-byte load covers the pixels a short edge crosses.
  int tx = (width + 7) / 8, ty = (height + 7) / 8;
  std::vector<unsigned long long> bits((size_t)tx * ty, 0ull);
  for (int r = 0; r < height; ++r)
    for (int c = 0; c < width; ++c)
      if ((int)pix[(size_t)r * width + c] >= threshold)
        bits[(size_t)(r >> 3) * tx + (c >> 3)] |= 1ull << (((r & 7) << 3) | (c & 7));
  BP_TRY(bp_reserve(ctx, ctx->d_bits, bits.size() * sizeof(unsigned long long), false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_bits.p, bits.data(), bits.size() * sizeof(unsigned long long),
                               cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->image.W = width;
  ctx->image.H = height;
  ctx->image.tiles_x = tx;
  ctx->image.bits = ctx->d_bits.as<unsigned long long>();
  ctx->world = WORLD_IMAGE;
  return BPOMP_OK;
}

extern "C" int bpomp_world_set_boxes(bpomp_ctx* ctx, const double* lo, const double* hi, int nboxes) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (nboxes < 0 || (nboxes > 0 && (!lo || !hi))) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad boxes");
  const int D = ctx->dim;
  BoxWorldView& w = ctx->boxes;
  memset(&w, 0, sizeof(w));
  w.nboxes = nboxes;
  int words = (nboxes + 31) / 32;
  int w4 = 1;
  while (w4 * 4 < words) w4 *= 2;
  w.w4 = w4;
  w.lohi_stride = D | 1;
  // Broad-phase grid: BP_GRID_C cells per dim over the bounding region of all boxes.
  for (int j = 0; j < D; ++j) {
    double mn = DBL_MAX, mx = -DBL_MAX;
    for (int b = 0; b < nboxes; ++b) {
      double l = lo[(size_t)b * D + j], h = hi[(size_t)b * D + j];
      if (l < mn) mn = l;
      if (h > mx) mx = h;
    }
    double span = mx - mn;
    w.g0[j] = (nboxes > 0 && std::isfinite(mn)) ? mn : 0.0;
    w.gs[j] = (nboxes > 0 && std::isfinite(span) && span > 0.0) ? (double)BP_GRID_C / span : 0.0;
    if (!std::isfinite(w.gs[j])) w.gs[j] = 0.0;
  }
  // Broad-phase masks (one bit per box).  bp_cell is monotone, so any x in [lo_j,hi_j] has cell(x) inside
  // the box's cell interval [cl,ch]; a segment whose endpoints span cells [c0,c1] in dimension j can only
  // touch boxes with cl <= c1 and ch >= c0.  Short ranges (c1-c0 < BP_SPAN) get one combined row
  // bp_rrow(j,c0,span); longer ones use two prefix rows, bp_row(j,0,c1) & bp_row(j,1,c0).  The AND over
  // all dimensions is a superset of the boxes the segment can touch.
  const int rows = D * BP_GRID_C * BP_SPAN;
  const int rows_ab = D * 2 * BP_GRID_C;
  w.rows = rows;
  w.rows_ab = rows_ab;
  w.ab_off = (size_t)rows * w4;
  std::vector<uint32_t> rmask(((size_t)rows + rows_ab) * w4 * 4, 0u);
  uint32_t* ab = rmask.data() + w.ab_off * 4;
  std::vector<double2> lohi((size_t)std::max(nboxes, 1) * w.lohi_stride, make_double2(1.0, 0.0));
  for (int b = 0; b < nboxes; ++b) {
    bool empty = false;
    for (int j = 0; j < D; ++j) {
      double l = lo[(size_t)b * D + j], h = hi[(size_t)b * D + j];
      lohi[(size_t)b * w.lohi_stride + j] = make_double2(l, h);
      if (!(l <= h)) empty = true;  // contains nothing (also NaN bounds)
    }
    if (empty) continue;
    const int word = b >> 5, g = word >> 2;
    const uint32_t bit = 1u << (b & 31);
    for (int j = 0; j < D; ++j) {
      int cl = bp_cell(lo[(size_t)b * D + j], w.g0[j], w.gs[j]);
      int ch = bp_cell(hi[(size_t)b * D + j], w.g0[j], w.gs[j]);
      for (int c = 0; c < BP_GRID_C; ++c) {
        if (cl <= c) ab[bp_midx(bp_row(j, 0, c), g, rows_ab) * 4 + (word & 3)] |= bit;
        if (ch >= c) ab[bp_midx(bp_row(j, 1, c), g, rows_ab) * 4 + (word & 3)] |= bit;
        for (int s = 0; s < BP_SPAN; ++s)
          if (cl <= c + s && ch >= c) rmask[bp_midx(bp_rrow(j, c, s), g, rows) * 4 + (word & 3)] |= bit;
      }
    }
  }
  w.rmask_bytes = (size_t)rows * w4 * 16;  // the part staged into shared memory (range masks only)
  w.lohi_bytes = (size_t)nboxes * w.lohi_stride * sizeof(double2);
  BP_TRY(bp_reserve(ctx, ctx->d_rmask, rmask.size() * sizeof(uint32_t), false));
  BP_TRY(bp_reserve(ctx, ctx->d_lohi, lohi.size() * sizeof(double2), false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_rmask.p, rmask.data(), rmask.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                               ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_lohi.p, lohi.data(), lohi.size() * sizeof(double2), cudaMemcpyHostToDevice,
                               ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  w.rmask = ctx->d_rmask.as<uint4>();
  w.lohi = ctx->d_lohi.as<double2>();
  ctx->world = WORLD_BOXES;
  ctx->h_box_lo.assign(lo, lo + (size_t)nboxes * D);
  ctx->h_box_hi.assign(hi, hi + (size_t)nboxes * D);
  return bp_build_fast_world(ctx);
}

// Tables of the fast kernel (fastpath.cuh): the quantisation q_j over the region of interest (the bounding
// region of the boxes, intersected with the state-space bounds when the caller gave them), the 16-bit box
// bounds, and the broad-phase masks in the team layout.  Worlds the fast kernel does not cover (more than
// FP_MAXBOXES boxes, non-finite or degenerate extents, coordinates huge relative to the extent) leave
// fast.enabled = 0 and are evaluated by the general kernel alone.
int bp_build_fast_world(bpomp_ctx* ctx) {
  FastWorldView& fw = ctx->fast;
  const float* perm_s = ctx->d_perm_s.as<float>();
  const float* perm4 = ctx->d_perm4.as<float>();
  memset(&fw, 0, sizeof(fw));
  fw.perm_s = perm_s;
  fw.perm4 = perm4;
  const int D = ctx->dim;
  const int nb = (int)(ctx->h_box_lo.size() / (size_t)D);
  fw.nboxes = nb;
  fw.nps = (D + 1) / 2;
  if (ctx->world != WORLD_BOXES) return BPOMP_OK;
  FpTables t = fp_build_tables(D, nb, ctx->h_box_lo.data(), ctx->h_box_hi.data(), ctx->have_bounds, ctx->bounds_lo,
                               ctx->bounds_hi);
  if (!t.enabled) return BPOMP_OK;
  for (int j = 0; j < D; ++j) {
    fw.qs[j] = t.qs[j];
    fw.qo[j] = t.qo[j];
  }
  fw.tmask_bytes = t.tmask.size() * sizeof(uint32_t);
  fw.qbox_bytes = t.qbox.size() * sizeof(uint16_t);
  BP_TRY(bp_reserve(ctx, ctx->d_tmask, fw.tmask_bytes, false));
  BP_TRY(bp_reserve(ctx, ctx->d_qbox, fw.qbox_bytes, false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_tmask.p, t.tmask.data(), fw.tmask_bytes, cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->d_qbox.p, t.qbox.data(), fw.qbox_bytes, cudaMemcpyHostToDevice, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  fw.tmask = ctx->d_tmask.as<uint4>();
  fw.qbox = ctx->d_qbox.as<uint4>();
  fw.enabled = 1;
  return BPOMP_OK;
}

extern "C" int bpomp_world_set_bounds(bpomp_ctx* ctx, const double* lo, const double* hi) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (!lo || !hi) {
    ctx->have_bounds = false;
  } else {
    for (int j = 0; j < ctx->dim; ++j) {
      if (!(lo[j] < hi[j]) || !std::isfinite(lo[j]) || !std::isfinite(hi[j]))
        return bp_fail(ctx, BPOMP_ERR_INVALID, "bounds must be finite with lo < hi in every dimension");
      ctx->bounds_lo[j] = lo[j];
      ctx->bounds_hi[j] = hi[j];
    }
    ctx->have_bounds = true;
  }
  if (ctx->world == WORLD_BOXES) return bp_build_fast_world(ctx);
  return BPOMP_OK;
}

// ---------------------------------------------------------------------------
// Batched isValid
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) states_valid_boxes_kernel(const double* __restrict__ states, int64_t count,
                                                                 BoxWorldView w, uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    double x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = __ldg(states + i * D + j);
    out[i] = bp_valid_boxes<D>(x, w, w.rmask, w.lohi) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) states_valid_image_kernel(const double* __restrict__ states, int64_t count,
                                                                 ImageWorldView im, uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    double2 x = __ldg(reinterpret_cast<const double2*>(states) + i);
    out[i] = bp_valid_image(x.x, x.y, im) ? 1 : 0;
  }
}

template <int D>
static void launch_valid_boxes(bpomp_ctx* ctx, const double* s, int64_t count, uint8_t* out, int grid) {
  states_valid_boxes_kernel<D><<<grid, 256, 0, ctx->stream>>>(s, count, ctx->boxes, out);
}

int bp_launch_states_valid(bpomp_ctx* ctx, const double* d_states, int64_t count, uint8_t* d_out) {
  if (ctx->world == WORLD_NONE) return bp_fail(ctx, BPOMP_ERR_NO_WORLD, "no collision world set");
  if (count <= 0) return BPOMP_OK;
  int grid = (int)std::min<int64_t>((count + 255) / 256, (int64_t)ctx->num_sms * 8);
  if (ctx->world == WORLD_IMAGE) {
    states_valid_image_kernel<<<grid, 256, 0, ctx->stream>>>(d_states, count, ctx->image, d_out);
  } else {
    switch (ctx->dim) {
      case 1: launch_valid_boxes<1>(ctx, d_states, count, d_out, grid); break;
      case 2: launch_valid_boxes<2>(ctx, d_states, count, d_out, grid); break;
      case 3: launch_valid_boxes<3>(ctx, d_states, count, d_out, grid); break;
      case 4: launch_valid_boxes<4>(ctx, d_states, count, d_out, grid); break;
      case 5: launch_valid_boxes<5>(ctx, d_states, count, d_out, grid); break;
      case 6: launch_valid_boxes<6>(ctx, d_states, count, d_out, grid); break;
      case 7: launch_valid_boxes<7>(ctx, d_states, count, d_out, grid); break;
      case 8: launch_valid_boxes<8>(ctx, d_states, count, d_out, grid); break;
      default: return bp_fail(ctx, BPOMP_ERR_INVALID, "unsupported dim %d", ctx->dim);
    }
  }
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  return BPOMP_OK;
}

extern "C" int bpomp_states_valid_dev(bpomp_ctx* ctx, const double* d_states, int64_t count, uint8_t* d_out) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!d_states || !d_out))) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  return bp_launch_states_valid(ctx, d_states, count, d_out);
}

extern "C" int bpomp_states_valid(bpomp_ctx* ctx, const double* states, int64_t count, uint8_t* out) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!states || !out))) return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (ctx->world == WORLD_NONE) return bp_fail(ctx, BPOMP_ERR_NO_WORLD, "no collision world set");
  if (count == 0) return BPOMP_OK;
  size_t bytes = (size_t)count * ctx->dim * sizeof(double);
  BP_TRY(bp_reserve(ctx, ctx->s_a, bytes, false));
  BP_TRY(bp_reserve(ctx, ctx->s_c, (size_t)count, false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_a.p, states, bytes, cudaMemcpyHostToDevice, ctx->stream));
  BP_TRY(bp_launch_states_valid(ctx, ctx->s_a.as<double>(), count, ctx->s_c.as<uint8_t>()));
  BP_CUDA(ctx, cudaMemcpyAsync(out, ctx->s_c.p, (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}

// ---------------------------------------------------------------------------
// vertexPruneFunction / isVertexInadmissible — BP.cpp:636-656
// ---------------------------------------------------------------------------
struct PruneArgs {
  double start[BPOMP_DIM_MAX], goal[BPOMP_DIM_MAX];
  double best;
  int use_cost;
};

template <int D>
__global__ void __launch_bounds__(256) states_prune_kernel(const double* __restrict__ states, int64_t count,
                                                           PruneArgs a, const uint8_t* __restrict__ valid,
                                                           uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    bool prune = false;
    if (a.use_cost) {  // BP.cpp:638-646
      double x[D];
#pragma unroll
      for (int j = 0; j < D; ++j) x[j] = __ldg(states + i * D + j);
      double ds = __dsqrt_rn(bp_dist2<D>(a.start, x));
      double dg = __dsqrt_rn(bp_dist2<D>(a.goal, x));
      prune = __dadd_rn(ds, dg) > a.best;
    }
    if (valid && !valid[i]) prune = true;  // BP.cpp:654
    out[i] = prune ? 1 : 0;
  }
}

extern "C" int bpomp_states_prune(bpomp_ctx* ctx, const double* states, int64_t count, const double* start,
                                  const double* goal, double best_cost, int check_validity, uint8_t* out) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  if (count < 0 || (count > 0 && (!states || !out)) || !start || !goal)
    return bp_fail(ctx, BPOMP_ERR_INVALID, "bad arguments");
  if (check_validity && ctx->world == WORLD_NONE) return bp_fail(ctx, BPOMP_ERR_NO_WORLD, "no collision world set");
  if (count == 0) return BPOMP_OK;
  const int D = ctx->dim;
  size_t bytes = (size_t)count * D * sizeof(double);
  BP_TRY(bp_reserve(ctx, ctx->s_a, bytes, false));
  BP_TRY(bp_reserve(ctx, ctx->s_c, (size_t)count, false));
  BP_TRY(bp_reserve(ctx, ctx->s_d, (size_t)count, false));
  BP_CUDA(ctx, cudaMemcpyAsync(ctx->s_a.p, states, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const uint8_t* d_valid = nullptr;
  if (check_validity) {
    BP_TRY(bp_launch_states_valid(ctx, ctx->s_a.as<double>(), count, ctx->s_c.as<uint8_t>()));
    d_valid = ctx->s_c.as<uint8_t>();
  }
  PruneArgs a;
  memset(&a, 0, sizeof(a));
  for (int j = 0; j < D; ++j) {
    a.start[j] = start[j];
    a.goal[j] = goal[j];
  }
  a.best = best_cost;
  a.use_cost = best_cost < DBL_MAX ? 1 : 0;  // BP.cpp:638
  int grid = (int)std::min<int64_t>((count + 255) / 256, (int64_t)ctx->num_sms * 8);
  const double* s = ctx->s_a.as<double>();
  uint8_t* o = ctx->s_d.as<uint8_t>();
  switch (D) {
    case 1: states_prune_kernel<1><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 2: states_prune_kernel<2><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 3: states_prune_kernel<3><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 4: states_prune_kernel<4><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 5: states_prune_kernel<5><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 6: states_prune_kernel<6><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 7: states_prune_kernel<7><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 8: states_prune_kernel<8><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    default: return bp_fail(ctx, BPOMP_ERR_INVALID, "unsupported dim %d", D);
  }
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  BP_CUDA(ctx, cudaMemcpyAsync(out, o, (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}

extern "C" int bpomp_vertices_prune(bpomp_ctx* ctx, uint32_t start_id, uint32_t goal_id, double best_cost,
                                    int check_validity, uint8_t* out) {
  if (!ctx) return BPOMP_ERR_INVALID;
  bp_enter(ctx);
  const uint32_t count = ctx->nverts;
  if (count == 0) return BPOMP_OK;
  if (!out) return bp_fail(ctx, BPOMP_ERR_INVALID, "out is NULL");
  if (start_id >= count || goal_id >= count) return bp_fail(ctx, BPOMP_ERR_RANGE, "start/goal id out of range");
  if (check_validity && ctx->world == WORLD_NONE) return bp_fail(ctx, BPOMP_ERR_NO_WORLD, "no collision world set");
  const int D = ctx->dim;
  BP_TRY(bp_reserve(ctx, ctx->s_c, (size_t)count, false));
  BP_TRY(bp_reserve(ctx, ctx->s_d, (size_t)count, false));
  const double* s = ctx->d_verts.as<double>();
  const uint8_t* d_valid = nullptr;
  if (check_validity) {
    BP_TRY(bp_launch_states_valid(ctx, s, count, ctx->s_c.as<uint8_t>()));
    d_valid = ctx->s_c.as<uint8_t>();
  }
  PruneArgs a;
  memset(&a, 0, sizeof(a));
  BP_CUDA(ctx, cudaMemcpyAsync(a.start, s + (size_t)start_id * D, D * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaMemcpyAsync(a.goal, s + (size_t)goal_id * D, D * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  a.best = best_cost;
  a.use_cost = best_cost < DBL_MAX ? 1 : 0;
  int grid = (int)std::min<int64_t>(((int64_t)count + 255) / 256, (int64_t)ctx->num_sms * 8);
  uint8_t* o = ctx->s_d.as<uint8_t>();
  switch (D) {
    case 1: states_prune_kernel<1><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 2: states_prune_kernel<2><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 3: states_prune_kernel<3><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 4: states_prune_kernel<4><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 5: states_prune_kernel<5><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 6: states_prune_kernel<6><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 7: states_prune_kernel<7><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    case 8: states_prune_kernel<8><<<grid, 256, 0, ctx->stream>>>(s, count, a, d_valid, o); break;
    default: return bp_fail(ctx, BPOMP_ERR_INVALID, "unsupported dim %d", D);
  }
  ctx->launches++;
  BP_CUDA(ctx, cudaGetLastError());
  BP_CUDA(ctx, cudaMemcpyAsync(out, o, (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  BP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return BPOMP_OK;
}
