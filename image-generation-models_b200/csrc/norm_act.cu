// GroupNorm(8)+Mish(+time-embedding add)(+residual) and the channel LayerNorm of the
// reference U-Net, forward and backward.  All HBM-bound: 128-bit NHWC loads,
// warp-shuffle reductions, deterministic (no float atomics on statistics).
//
// Replaces: nn.GroupNorm + Mish in Block (reference src/models/ddpm.py:116, :62-64),
// the `h += mlp(time_emb)[:, :, None, None]` add (:140), the residual add (:143),
// LayerNorm (:85-95) and their autograd.
#include "common.cuh"

#include <algorithm>
#include <cooperative_groups.h>
#include <stdlib.h>

namespace igm {
namespace {

// Thread layout shared by the GroupNorm kernels: L = C/4 lanes per pixel (one
// float4 of channels each), PPI = 256/L pixels in flight per iteration.
// Requires C % 32 == 0 (so a float4 never straddles a group) and C <= 1024.
struct GnLayout {
  int L, PPI, c4, pslot, cpg, lpg, group;
  __device__ GnLayout(int C) {
    L = C >> 2;
    cpg = C >> 3;      // kGroups == 8
    lpg = cpg >> 2;
    if ((L & (L - 1)) == 0) {   // power of two (every width of the reference's configs): shifts, no integer division
      const int lg = 31 - __clz(L);
      PPI = 256 >> lg;
      c4 = threadIdx.x & (L - 1);
      pslot = threadIdx.x >> lg;
      group = c4 >> (lg - 3);
    } else {
      PPI = 256 / L;
      c4 = threadIdx.x % L;
      pslot = threadIdx.x / L;
      group = c4 / lpg;
    }
  }
};

// Deterministic per-group reduction of one (a, b) pair per thread.
// Exactly 32 threads belong to each group (lpg * PPI == 32): warp g reduces group g.
__device__ __forceinline__ void group_reduce2(const GnLayout& ly, float a, float b, float (*sm)[2],
                                              float& ra, float& rb) {
  sm[threadIdx.x][0] = a;
  sm[threadIdx.x][1] = b;
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int src = (lane / ly.lpg) * ly.L + w * ly.lpg + (lane % ly.lpg);
  ra = warp_sum(sm[src][0]);
  rb = warp_sum(sm[src][1]);
  __syncthreads();
}

__global__ void __launch_bounds__(256) gn_partial_kernel(const float* __restrict__ y, int HW, int C,
                                                         int nchunks, float* __restrict__ part, int kGnChunk) {
  __shared__ float sm[256][2];
  const GnLayout ly(C);
  const int b = blockIdx.x / nchunks, chunk = blockIdx.x % nchunks;
  const int p0 = chunk * kGnChunk;
  const int p1 = min(p0 + kGnChunk, HW);
  float s = 0.f, ss = 0.f;
  const float* base = y + ((int64_t)b * HW) * C + ly.c4 * 4;
  for (int p = p0 + ly.pslot; p < p1; p += ly.PPI) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)p * C));
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  float rs, rss;
  group_reduce2(ly, s, ss, sm, rs, rss);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    float* o = part + (((int64_t)b * nchunks + chunk) * kGroups + w) * 2;
    o[0] = rs;
    o[1] = rss;
  }
}

// mean / rstd of sample b, group g from the chunk partials (fp64 combine)
__device__ __forceinline__ void finalize_stats(const float* part, int b, int nchunks, int g, int count,
                                               float& mean, float& rstd) {
  double s = 0.0, ss = 0.0;
  for (int c = 0; c < nchunks; ++c) {
    const float* o = part + (((int64_t)b * nchunks + c) * kGroups + g) * 2;
    s += (double)o[0];
    ss += (double)o[1];
  }
  const double m = s / count;
  double var = ss / count - m * m;
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)kGnEps));
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ y, const float* __restrict__ part,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta,
                                                       const float* __restrict__ temb, int temb_stride,
                                                       const float* __restrict__ res, float* __restrict__ out,
                                                       float* __restrict__ stats, int HW, int C, int nchunks,
                                                       __nv_bfloat16* __restrict__ out_hi,
                                                       __nv_bfloat16* __restrict__ out_lo, int nparts, int kGnChunk) {
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  const GnLayout ly(C);
  const int b = blockIdx.x / nchunks, chunk = blockIdx.x % nchunks;
  pdl_wait();
  if (threadIdx.x < kGroups) {
    float m, r;
    finalize_stats(part, b, nparts, threadIdx.x, HW * ly.cpg, m, r);
    s_mean[threadIdx.x] = m;
    s_rstd[threadIdx.x] = r;
    if (chunk == 0 && stats) {
      stats[((int64_t)b * kGroups + threadIdx.x) * 2 + 0] = m;
      stats[((int64_t)b * kGroups + threadIdx.x) * 2 + 1] = r;
    }
  }
  __syncthreads();
  const float mean = s_mean[ly.group], rstd = s_rstd[ly.group];
  const int c = ly.c4 * 4;
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
  float4 te = make_float4(0.f, 0.f, 0.f, 0.f);
  if (temb) te = __ldg(reinterpret_cast<const float4*>(temb + (int64_t)b * temb_stride + c));
  const int p0 = chunk * kGnChunk;
  const int p1 = min(p0 + kGnChunk, HW);
  for (int p = p0 + ly.pslot; p < p1; p += ly.PPI) {
    const int64_t off = ((int64_t)b * HW + p) * C + c;
    const float4 v = __ldg(reinterpret_cast<const float4*>(y + off));
    float4 o;
    o.x = mish_f((v.x - mean) * rstd * ga.x + be.x) + te.x;
    o.y = mish_f((v.y - mean) * rstd * ga.y + be.y) + te.y;
    o.z = mish_f((v.z - mean) * rstd * ga.z + be.z) + te.z;
    o.w = mish_f((v.w - mean) * rstd * ga.w + be.w) + te.w;
    if (res) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(res + off));
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (out) *reinterpret_cast<float4*>(out + off) = o;
    if (out_hi) store_split4(out_hi, out_lo, off, o);
  }
}

// ---- forward apply, fast path: C/4 a power of two -------------------------------------------------------------------
// The generic kernel above was bound by instruction issue and by its prologue, not by HBM (ncu: 61 % issue-active,
// barrier the top stall, DRAM 16 % busy): eight threads folded the HW/32 statistics slots serially in fp64 while 248
// waited, index arithmetic used integer divisions, and Mish / the bf16 split carried range-handling code.  Here the
// thread layout is shifts and masks, all four pixel loads of a thread are in flight before the statistics are touched,
// warp g folds group g's slots with one load round and five shuffles (fp32 tree, fixed order; mean / variance arithmetic
// in fp64 as before), Mish is ex2.approx + rcp.approx with a select, and the split packs two values per cvt.
constexpr int kGnIter = 4;   // float4 per thread

__global__ void __launch_bounds__(256) gn_apply_fast_kernel(const float* __restrict__ y, const float* __restrict__ part,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ temb, int temb_stride,
                                                            const float* __restrict__ res, float* __restrict__ out,
                                                            float* __restrict__ stats, int HW, int C, int lgL,
                                                            __nv_bfloat16* __restrict__ out_hi,
                                                            __nv_bfloat16* __restrict__ out_lo, int nparts, double inv_count) {
  __shared__ float s_mean[kGroups], s_rstd[kGroups];
  const int tid = threadIdx.x, b = blockIdx.y;
  const int c4 = tid & ((1 << lgL) - 1), pslot = tid >> lgL, ppi = 256 >> lgL;
  const int group = c4 >> (lgL - 3);
  const int c = c4 * 4;
  // parameters and the time-embedding projection were written at least two kernels back: load them before the wait
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
  float4 te = make_float4(0.f, 0.f, 0.f, 0.f);
  if (temb) te = __ldg(reinterpret_cast<const float4*>(temb + (int64_t)b * temb_stride + c));
  const int p0 = blockIdx.x * (ppi * kGnIter) + pslot;
  const int64_t off0 = ((int64_t)b * HW + p0) * C + c;
  const int64_t step = (int64_t)ppi * C;
  pdl_wait();
  float4 v[kGnIter], r[kGnIter];
#pragma unroll
  for (int i = 0; i < kGnIter; ++i) {
    const bool ok = p0 + i * ppi < HW;
    v[i] = ok ? __ldg(reinterpret_cast<const float4*>(y + off0 + i * step)) : make_float4(0.f, 0.f, 0.f, 0.f);
    r[i] = (ok && res) ? __ldg(reinterpret_cast<const float4*>(res + off0 + i * step)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  {
    const int g = tid >> 5, lane = tid & 31;
    float s = 0.f, ss = 0.f;
    for (int k = lane; k < nparts; k += 32) {
      const float2 o = __ldg(reinterpret_cast<const float2*>(part + (((int64_t)b * nparts + k) * kGroups + g) * 2));
      s += o.x;
      ss += o.y;
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane == 0) {
      const double m = (double)s * inv_count;
      double var = (double)ss * inv_count - m * m;
      if (var < 0.0) var = 0.0;
      const float mean = (float)m, rstd = rsqrtf((float)var + kGnEps);
      s_mean[g] = mean;
      s_rstd[g] = rstd;
      if (blockIdx.x == 0 && stats) {
        stats[((int64_t)b * kGroups + g) * 2 + 0] = mean;
        stats[((int64_t)b * kGroups + g) * 2 + 1] = rstd;
      }
    }
  }
  __syncthreads();
  const float mean = s_mean[group], rstd = s_rstd[group];
  // (x - mean) * rstd * gamma + beta  ==  x * a + d
  const float4 sa = make_float4(rstd * ga.x, rstd * ga.y, rstd * ga.z, rstd * ga.w);
  const float4 sd = make_float4(be.x - mean * sa.x, be.y - mean * sa.y, be.z - mean * sa.z, be.w - mean * sa.w);
#pragma unroll
  for (int i = 0; i < kGnIter; ++i) {
    if (p0 + i * ppi >= HW) break;
    float4 o;
    o.x = mish_f(fmaf(v[i].x, sa.x, sd.x)) + te.x + r[i].x;
    o.y = mish_f(fmaf(v[i].y, sa.y, sd.y)) + te.y + r[i].y;
    o.z = mish_f(fmaf(v[i].z, sa.z, sd.z)) + te.z + r[i].z;
    o.w = mish_f(fmaf(v[i].w, sa.w, sd.w)) + te.w + r[i].w;
    const int64_t off = off0 + i * step;
    if (out) *reinterpret_cast<float4*>(out + off) = o;
    if (out_hi) {
      uint2 h, l;
      split_pair(o.x, o.y, h.x, l.x);
      split_pair(o.z, o.w, h.y, l.y);
      *reinterpret_cast<uint2*>(out_hi + off) = h;
      *reinterpret_cast<uint2*>(out_lo + off) = l;
    }
  }
}

// ---- backward, pass 1: per-chunk reductions ---------------------------------
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const GnBwdArgs a, int nchunks, int kGnChunk) {
  __shared__ float sm[256][2];
  __shared__ float4 sc[256][3];
  const GnLayout ly(a.C);
  const int b = blockIdx.x / nchunks, chunk = blockIdx.x % nchunks;
  const float mean = __ldg(a.stats + ((int64_t)b * kGroups + ly.group) * 2 + 0);
  const float rstd = __ldg(a.stats + ((int64_t)b * kGroups + ly.group) * 2 + 1);
  const int c = ly.c4 * 4;
  const float4 ga4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
  const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w};
  const float be[4] = {be4.x, be4.y, be4.z, be4.w};
  float dgam[4] = {0.f, 0.f, 0.f, 0.f}, dbet[4] = {0.f, 0.f, 0.f, 0.f}, dte[4] = {0.f, 0.f, 0.f, 0.f};
  float s1 = 0.f, s2 = 0.f;
  const int p0 = chunk * kGnChunk;
  const int p1 = min(p0 + kGnChunk, a.HW);
  for (int p = p0 + ly.pslot; p < p1; p += ly.PPI) {
    const int64_t off = ((int64_t)b * a.HW + p) * a.C + c;
    const float4 y4 = __ldg(reinterpret_cast<const float4*>(a.y + off));
    const float4 d4 = __ldg(reinterpret_cast<const float4*>(a.d_out + off));
    const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
    const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float n = (yv[j] - mean) * rstd;
      const float g = n * ga[j] + be[j];
      const float dg = dv[j] * mish_grad_f(g);
      dgam[j] += dg * n;
      dbet[j] += dg;
      dte[j] += dv[j];
      const float dn = dg * ga[j];
      s1 += dn;
      s2 += dn * n;
    }
  }
  float r1, r2;
  group_reduce2(ly, s1, s2, sm, r1, r2);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    float* o = a.ws_group + (((int64_t)b * nchunks + chunk) * kGroups + w) * 2;
    o[0] = r1;
    o[1] = r2;
  }
  // per-channel partials: reduce over the PPI pixel slots in a fixed order
  sc[threadIdx.x][0] = make_float4(dgam[0], dgam[1], dgam[2], dgam[3]);
  sc[threadIdx.x][1] = make_float4(dbet[0], dbet[1], dbet[2], dbet[3]);
  sc[threadIdx.x][2] = make_float4(dte[0], dte[1], dte[2], dte[3]);
  __syncthreads();
  if (threadIdx.x < ly.L) {
    float4 acc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < ly.PPI; ++s) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float4 v = sc[s * ly.L + threadIdx.x][k];
        acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
      }
    }
    // ws_chan layout: [b][chunk][4][C]  (dgamma, dbeta, dtemb partials; slot 3 = conv-bias partial of pass 2)
    float* o = a.ws_chan + ((int64_t)b * nchunks + chunk) * 4 * a.C + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 3; ++k) *reinterpret_cast<float4*>(o + (int64_t)k * a.C) = acc[k];
  }
}

// ---- backward, pass 2: dy ------------------------------------------------------
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const GnBwdArgs a, int nchunks, int kGnChunk) {
  __shared__ float s_m1[kGroups], s_m2[kGroups];
  __shared__ float4 s_db[256];
  float dbs[4] = {0.f, 0.f, 0.f, 0.f};
  const GnLayout ly(a.C);
  const int b = blockIdx.x / nchunks, chunk = blockIdx.x % nchunks;
  if (threadIdx.x < kGroups) {
    double s1 = 0.0, s2 = 0.0;
    for (int c = 0; c < nchunks; ++c) {
      const float* o = a.ws_group + (((int64_t)b * nchunks + c) * kGroups + threadIdx.x) * 2;
      s1 += (double)o[0];
      s2 += (double)o[1];
    }
    const double cnt = (double)a.HW * ly.cpg;
    s_m1[threadIdx.x] = (float)(s1 / cnt);
    s_m2[threadIdx.x] = (float)(s2 / cnt);
  }
  __syncthreads();
  const float mean = __ldg(a.stats + ((int64_t)b * kGroups + ly.group) * 2 + 0);
  const float rstd = __ldg(a.stats + ((int64_t)b * kGroups + ly.group) * 2 + 1);
  const float m1 = s_m1[ly.group], m2 = s_m2[ly.group];
  const int c = ly.c4 * 4;
  const float4 ga4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
  const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w};
  const float be[4] = {be4.x, be4.y, be4.z, be4.w};
  const int p0 = chunk * kGnChunk;
  const int p1 = min(p0 + kGnChunk, a.HW);
  for (int p = p0 + ly.pslot; p < p1; p += ly.PPI) {
    const int64_t off = ((int64_t)b * a.HW + p) * a.C + c;
    const float4 y4 = __ldg(reinterpret_cast<const float4*>(a.y + off));
    const float4 d4 = __ldg(reinterpret_cast<const float4*>(a.d_out + off));
    const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
    const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float n = (yv[j] - mean) * rstd;
      const float g = n * ga[j] + be[j];
      const float dn = dv[j] * mish_grad_f(g) * ga[j];
      o[j] = rstd * (dn - m1 - n * m2);
      dbs[j] += o[j];
    }
    if (a.dy) *reinterpret_cast<float4*>(a.dy + off) = make_float4(o[0], o[1], o[2], o[3]);
    if (a.dy_hi) store_split4(a.dy_hi, a.dy_lo, off, make_float4(o[0], o[1], o[2], o[3]));
  }
  if (a.dbias) {
    // conv-bias gradient = column sums of dy: fold the pixel slots; the per-CTA partial goes to slot 3 of
    // ws_chan and is summed by pass 3 (same-address atomics from thousands of CTAs serialise)
    s_db[threadIdx.x] = make_float4(dbs[0], dbs[1], dbs[2], dbs[3]);
    __syncthreads();
    if (threadIdx.x < ly.L) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int sidx = 0; sidx < ly.PPI; ++sidx) {
        const float4 v = s_db[sidx * ly.L + threadIdx.x];
        t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
      }
      *reinterpret_cast<float4*>(a.ws_chan + (((int64_t)b * nchunks + chunk) * 4 + 3) * a.C + threadIdx.x * 4) = t;
    }
  }
}
// ---- backward, fused: one thread-block cluster per sample ----------------------------------------
// The two-pass scheme above reads y and d_out twice and needs three launches.  When a sample splits into
// CS <= 8 CTAs of exactly 8192 elements, each thread keeps its NV = 8 float4 of (n, dn) in registers, the
// group sums cross the cluster through distributed shared memory, and dy / the parameter gradients come out of
// the same kernel: y and d_out are read once.  NV = float4 per thread (8 unless the tensor is too small).
template <int GN_NV>
__global__ void __launch_bounds__(256) gn_bwd_fused_kernel(const GnBwdArgs a, int cs) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float sm[256][2];
  __shared__ float cl_grp[kGroups][2];    // this CTA's (sum dn, sum dn*n) per group, read by the whole cluster
  __shared__ float s_m[kGroups][2];
  __shared__ float4 sc[256][4];           // per-thread channel partials -> per-channel CTA totals in sc[0..L)
  pdl_wait();
  const GnLayout ly(a.C);
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / cs;
  const int npix = a.HW / cs;             // pixels of this CTA: NV * PPI
  const int p0 = rank * npix;
  const float mean = __ldg(a.stats + ((int64_t)b * kGroups + ly.group) * 2 + 0);
  const float rstd = __ldg(a.stats + ((int64_t)b * kGroups + ly.group) * 2 + 1);
  const int c = ly.c4 * 4;
  const float4 ga4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
  const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w};
  const float be[4] = {be4.x, be4.y, be4.z, be4.w};
  float4 nv[GN_NV], dnv[GN_NV];
  const int64_t base = ((int64_t)b * a.HW + p0 + ly.pslot) * a.C + c;
  const int64_t step = (int64_t)ly.PPI * a.C;
#pragma unroll
  for (int i = 0; i < GN_NV; ++i) {
    nv[i] = __ldg(reinterpret_cast<const float4*>(a.y + base + i * step));
    dnv[i] = __ldg(reinterpret_cast<const float4*>(a.d_out + base + i * step));
  }
  if (a.dout_hi) {
#pragma unroll
    for (int i = 0; i < GN_NV; ++i) store_split4(a.dout_hi, a.dout_lo, base + i * step, dnv[i]);
  }
  float dgam[4] = {0.f, 0.f, 0.f, 0.f}, dbet[4] = {0.f, 0.f, 0.f, 0.f}, dte[4] = {0.f, 0.f, 0.f, 0.f};
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < GN_NV; ++i) {
    float* yv = reinterpret_cast<float*>(&nv[i]);
    float* dv = reinterpret_cast<float*>(&dnv[i]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float n = (yv[j] - mean) * rstd;
      const float g = n * ga[j] + be[j];
      const float dg = dv[j] * mish_grad_f(g);
      dgam[j] += dg * n;
      dbet[j] += dg;
      dte[j] += dv[j];
      const float dn = dg * ga[j];
      s1 += dn;
      s2 += dn * n;
      yv[j] = n;
      dv[j] = dn;
    }
  }
  float r1, r2;
  group_reduce2(ly, s1, s2, sm, r1, r2);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { cl_grp[w][0] = r1; cl_grp[w][1] = r2; }
  cluster.sync();
  if (threadIdx.x < kGroups * 2) {
    const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
    float t = 0.f;
    for (int r = 0; r < cs; ++r) t += cluster.map_shared_rank(&cl_grp[0][0], r)[g * 2 + k];   // fixed order
    s_m[g][k] = t / ((float)a.HW * ly.cpg);
  }
  __syncthreads();
  const float m1 = s_m[ly.group][0], m2 = s_m[ly.group][1];
  float dbs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < GN_NV; ++i) {
    const float* n = reinterpret_cast<const float*>(&nv[i]);
    const float* dn = reinterpret_cast<const float*>(&dnv[i]);
    float4 o;
    o.x = rstd * (dn[0] - m1 - n[0] * m2);
    o.y = rstd * (dn[1] - m1 - n[1] * m2);
    o.z = rstd * (dn[2] - m1 - n[2] * m2);
    o.w = rstd * (dn[3] - m1 - n[3] * m2);
    dbs[0] += o.x; dbs[1] += o.y; dbs[2] += o.z; dbs[3] += o.w;
    const int64_t off = base + i * step;
    if (a.dy) *reinterpret_cast<float4*>(a.dy + off) = o;
    if (a.dy_hi) store_split4(a.dy_hi, a.dy_lo, off, o);
  }
  // per-channel sums: fold the pixel slots of the CTA, then the CTAs of the cluster (rank 0), then one atomic per
  // (sample, channel) -- dtemb is per sample and written directly
  sc[threadIdx.x][0] = make_float4(dgam[0], dgam[1], dgam[2], dgam[3]);
  sc[threadIdx.x][1] = make_float4(dbet[0], dbet[1], dbet[2], dbet[3]);
  sc[threadIdx.x][2] = make_float4(dte[0], dte[1], dte[2], dte[3]);
  sc[threadIdx.x][3] = make_float4(dbs[0], dbs[1], dbs[2], dbs[3]);
  __syncthreads();
  float4 acc[4];
  if (threadIdx.x < ly.L) {
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = sc[threadIdx.x][k];
    for (int sl = 1; sl < ly.PPI; ++sl) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 v = sc[sl * ly.L + threadIdx.x][k];
        acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < ly.L) {
#pragma unroll
    for (int k = 0; k < 4; ++k) sc[threadIdx.x][k] = acc[k];
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x < ly.L) {
    for (int r = 1; r < cs; ++r) {
      const float4* rs = cluster.map_shared_rank(&sc[0][0], r) + threadIdx.x * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 v = rs[k];
        acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
      }
    }
    atomicAdd(a.dgamma + c + 0, acc[0].x); atomicAdd(a.dgamma + c + 1, acc[0].y);
    atomicAdd(a.dgamma + c + 2, acc[0].z); atomicAdd(a.dgamma + c + 3, acc[0].w);
    atomicAdd(a.dbeta + c + 0, acc[1].x); atomicAdd(a.dbeta + c + 1, acc[1].y);
    atomicAdd(a.dbeta + c + 2, acc[1].z); atomicAdd(a.dbeta + c + 3, acc[1].w);
    if (a.dtemb) *reinterpret_cast<float4*>(a.dtemb + (int64_t)b * a.dtemb_stride + c) = acc[2];
    if (a.dout_colsum) {
      atomicAdd(a.dout_colsum + c + 0, acc[2].x); atomicAdd(a.dout_colsum + c + 1, acc[2].y);
      atomicAdd(a.dout_colsum + c + 2, acc[2].z); atomicAdd(a.dout_colsum + c + 3, acc[2].w);
    }
    if (a.dbias) {
      atomicAdd(a.dbias + c + 0, acc[3].x); atomicAdd(a.dbias + c + 1, acc[3].y);
      atomicAdd(a.dbias + c + 2, acc[3].z); atomicAdd(a.dbias + c + 3, acc[3].w);
    }
  }
  cluster.sync();   // remote shared memory stays valid until rank 0 has read it
}

// ---- backward, fused, bulk-staged variant (default) -------------------------------------------------------------
// Same decomposition as gn_bwd_fused_kernel (one cluster per sample, group sums through distributed shared memory), but a
// CTA's slice of y and d_out -- E = nv * 1024 consecutive floats each, NHWC pixels are contiguous -- is brought in by TWO
// 1-D bulk async copies (cp.async.bulk, one elected thread, completion on an mbarrier) into shared memory instead of 2 * nv
// float4 registers per thread.  The register version needed 128 registers (two CTAs per SM) and ran its load / reduce /
// cluster-barrier / store phases in lock step: ncu showed it barrier-stalled at 22 % occupancy and 31 % of the HBM rate.
// Here a thread keeps ~50 registers, three CTAs fit an SM (64 KB tiles), and the copies of one CTA overlap the compute
// and barrier phases of its neighbours.  n and dn are written back to the thread's own tile slots after pass 1, so pass 2
// does not recompute the Mish derivative.
__device__ __forceinline__ void bulk_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst_smem)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}

__global__ void __launch_bounds__(256) gn_bwd_bulk_kernel(const GnBwdArgs a, int cs, int nv) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(128) float tile[];   // y slice [E] | d_out slice [E]; later the per-thread channel partials
  __shared__ float sm[256][2];
  __shared__ float cl_grp[kGroups][2];    // this CTA's (sum dn, sum dn*n) per group, read by the whole cluster
  __shared__ float s_m[kGroups][2];
  __shared__ __align__(8) uint64_t bar;
  const int E = nv * 1024;
  float* ty = tile;
  float* td = tile + E;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();
  const GnLayout ly(a.C);
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / cs;
  const int npix = a.HW / cs;             // pixels of this CTA: nv * PPI
  const int p0 = rank * npix;
  const int64_t cta_base = ((int64_t)b * a.HW + p0) * a.C;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)),
                 "r"((uint32_t)(2 * E * sizeof(float)))
                 : "memory");
    bulk_load_1d(ty, a.y + cta_base, (uint32_t)(E * sizeof(float)), &bar);
    bulk_load_1d(td, a.d_out + cta_base, (uint32_t)(E * sizeof(float)), &bar);
  }
  const float mean = __ldg(a.stats + ((int64_t)b * kGroups + ly.group) * 2 + 0);
  const float rstd = __ldg(a.stats + ((int64_t)b * kGroups + ly.group) * 2 + 1);
  const float nmr = -mean * rstd;
  const int c = ly.c4 * 4;
  const float4 ga4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
  const float4 be4 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
  const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w};
  const float be[4] = {be4.x, be4.y, be4.z, be4.w};
  const int loc0 = ly.pslot * a.C + c;    // this thread's first float inside the tile
  const int lstep = ly.PPI * a.C;
  const int64_t base = cta_base + loc0;
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"((uint32_t)__cvta_generic_to_shared(&bar))
          : "memory");
    }
  }
  float dgam[4] = {0.f, 0.f, 0.f, 0.f}, dbet[4] = {0.f, 0.f, 0.f, 0.f}, dte[4] = {0.f, 0.f, 0.f, 0.f};
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 2
  for (int i = 0; i < nv; ++i) {
    float4 y4 = *reinterpret_cast<const float4*>(ty + loc0 + i * lstep);
    float4 d4 = *reinterpret_cast<const float4*>(td + loc0 + i * lstep);
    if (a.dout_hi) store_split4(a.dout_hi, a.dout_lo, base + (int64_t)i * lstep, d4);
    float* yv = reinterpret_cast<float*>(&y4);
    float* dv = reinterpret_cast<float*>(&d4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float n = fmaf(yv[j], rstd, nmr);   // (y - mean) * rstd
      const float g = fmaf(n, ga[j], be[j]);
      const float dg = dv[j] * mish_grad_nb_f(g);
      dgam[j] += dg * n;
      dbet[j] += dg;
      dte[j] += dv[j];
      const float dn = dg * ga[j];
      s1 += dn;
      s2 += dn * n;
      yv[j] = n;
      dv[j] = dn;
    }
    *reinterpret_cast<float4*>(ty + loc0 + i * lstep) = y4;   // own slots only: no barrier needed before pass 2
    *reinterpret_cast<float4*>(td + loc0 + i * lstep) = d4;
  }
  float r1, r2;
  group_reduce2(ly, s1, s2, sm, r1, r2);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { cl_grp[w][0] = r1; cl_grp[w][1] = r2; }
  cluster.sync();
  if (threadIdx.x < kGroups * 2) {
    const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
    float t = 0.f;
    for (int r = 0; r < cs; ++r) t += cluster.map_shared_rank(&cl_grp[0][0], r)[g * 2 + k];   // fixed order
    s_m[g][k] = t / ((float)a.HW * ly.cpg);
  }
  __syncthreads();
  const float m1 = s_m[ly.group][0], m2 = s_m[ly.group][1];
  const float q1 = -m1 * rstd, q2 = -m2 * rstd;
  float dbs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
  for (int i = 0; i < nv; ++i) {
    const float4 n = *reinterpret_cast<const float4*>(ty + loc0 + i * lstep);
    const float4 dn = *reinterpret_cast<const float4*>(td + loc0 + i * lstep);
    float4 o;
    o.x = fmaf(n.x, q2, fmaf(dn.x, rstd, q1));   // rstd * (dn - m1 - n * m2)
    o.y = fmaf(n.y, q2, fmaf(dn.y, rstd, q1));
    o.z = fmaf(n.z, q2, fmaf(dn.z, rstd, q1));
    o.w = fmaf(n.w, q2, fmaf(dn.w, rstd, q1));
    dbs[0] += o.x; dbs[1] += o.y; dbs[2] += o.z; dbs[3] += o.w;
    const int64_t off = base + (int64_t)i * lstep;
    if (a.dy) *reinterpret_cast<float4*>(a.dy + off) = o;
    if (a.dy_hi) store_split4(a.dy_hi, a.dy_lo, off, o);
  }
  // per-channel sums: fold the pixel slots of the CTA, then the CTAs of the cluster (rank 0), then one atomic per
  // (sample, channel) -- dtemb is per sample and written directly.  The tile is dead now: it holds the partials.
  __syncthreads();
  float4 (*sc)[4] = reinterpret_cast<float4 (*)[4]>(tile);
  sc[threadIdx.x][0] = make_float4(dgam[0], dgam[1], dgam[2], dgam[3]);
  sc[threadIdx.x][1] = make_float4(dbet[0], dbet[1], dbet[2], dbet[3]);
  sc[threadIdx.x][2] = make_float4(dte[0], dte[1], dte[2], dte[3]);
  sc[threadIdx.x][3] = make_float4(dbs[0], dbs[1], dbs[2], dbs[3]);
  __syncthreads();
  float4 acc[4];
  if (threadIdx.x < ly.L) {
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = sc[threadIdx.x][k];
    for (int sl = 1; sl < ly.PPI; ++sl) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 v = sc[sl * ly.L + threadIdx.x][k];
        acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < ly.L) {
#pragma unroll
    for (int k = 0; k < 4; ++k) sc[threadIdx.x][k] = acc[k];
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x < ly.L) {
    for (int r = 1; r < cs; ++r) {
      const float4* rs = cluster.map_shared_rank(&sc[0][0], r) + threadIdx.x * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 v = rs[k];
        acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w;
      }
    }
    atomicAdd(a.dgamma + c + 0, acc[0].x); atomicAdd(a.dgamma + c + 1, acc[0].y);
    atomicAdd(a.dgamma + c + 2, acc[0].z); atomicAdd(a.dgamma + c + 3, acc[0].w);
    atomicAdd(a.dbeta + c + 0, acc[1].x); atomicAdd(a.dbeta + c + 1, acc[1].y);
    atomicAdd(a.dbeta + c + 2, acc[1].z); atomicAdd(a.dbeta + c + 3, acc[1].w);
    if (a.dtemb) *reinterpret_cast<float4*>(a.dtemb + (int64_t)b * a.dtemb_stride + c) = acc[2];
    if (a.dout_colsum) {
      atomicAdd(a.dout_colsum + c + 0, acc[2].x); atomicAdd(a.dout_colsum + c + 1, acc[2].y);
      atomicAdd(a.dout_colsum + c + 2, acc[2].z); atomicAdd(a.dout_colsum + c + 3, acc[2].w);
    }
    if (a.dbias) {
      atomicAdd(a.dbias + c + 0, acc[3].x); atomicAdd(a.dbias + c + 1, acc[3].y);
      atomicAdd(a.dbias + c + 2, acc[3].z); atomicAdd(a.dbias + c + 3, acc[3].w);
    }
  }
  cluster.sync();   // remote shared memory stays valid until rank 0 has read it
}

// ---- backward, pass 3: parameter / time-embedding gradients -------------------
// grid (C/32 column slabs, B samples), block (32, 8): each CTA folds the chunk partials of one
// sample; dtemb[b, c] is written directly, dgamma/dbeta get one atomic per (sample, channel).
__global__ void __launch_bounds__(256) gn_bwd_param_kernel(const GnBwdArgs a, int nchunks) {
  __shared__ float red[8][4][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.y;
  float sg = 0.f, sb = 0.f, st = 0.f, sc = 0.f;
  if (c < a.C) {
    for (int ch = threadIdx.y; ch < nchunks; ch += 8) {
      const float* o = a.ws_chan + ((int64_t)b * nchunks + ch) * 4 * a.C;
      sg += o[c];
      sb += o[a.C + c];
      st += o[2 * a.C + c];
      if (a.dbias) sc += o[3 * a.C + c];
    }
  }
  red[threadIdx.y][0][threadIdx.x] = sg;
  red[threadIdx.y][1][threadIdx.x] = sb;
  red[threadIdx.y][2][threadIdx.x] = st;
  red[threadIdx.y][3][threadIdx.x] = sc;
  __syncthreads();
  if (threadIdx.y == 0 && c < a.C) {
    float tg = 0.f, tb = 0.f, tt = 0.f, tc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      tg += red[i][0][threadIdx.x];
      tb += red[i][1][threadIdx.x];
      tt += red[i][2][threadIdx.x];
      tc += red[i][3][threadIdx.x];
    }
    atomicAdd(a.dgamma + c, tg);
    atomicAdd(a.dbeta + c, tb);
    if (a.dbias) atomicAdd(a.dbias + c, tc);
    if (a.dtemb) a.dtemb[(int64_t)b * a.dtemb_stride + c] = tt;
  }
}

// ---------------------------------------------------------------------------
// LayerNorm over channels, one warp per pixel
// ---------------------------------------------------------------------------
constexpr int LN_MAX_V = 8;   // C <= 1024

__global__ void __launch_bounds__(256) ln_forward_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                         const float* __restrict__ bta, float* __restrict__ out,
                                                         int64_t M, int C, __nv_bfloat16* __restrict__ out_hi,
                                                         __nv_bfloat16* __restrict__ out_lo) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nv = C >> 2;
  for (int64_t m = warp; m < M; m += nwarps) {
    float4 v[LN_MAX_V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_V; ++j) {
      const int c4 = lane + j * 32;
      if (c4 < nv) {
        v[j] = __ldg(reinterpret_cast<const float4*>(x + m * C + c4 * 4));
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_V; ++j) {
      const int c4 = lane + j * 32;
      if (c4 < nv) {
        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
        q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
      }
    }
    const float stdv = sqrtf(warp_sum(q) / C);
    const float inv = 1.f / (stdv + kLnEps);
#pragma unroll
    for (int j = 0; j < LN_MAX_V; ++j) {
      const int c4 = lane + j * 32;
      if (c4 < nv) {
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c4 * 4));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bta + c4 * 4));
        float4 o;
        o.x = v[j].x * inv * gg.x + bb.x;
        o.y = v[j].y * inv * gg.y + bb.y;
        o.z = v[j].z * inv * gg.z + bb.z;
        o.w = v[j].w * inv * gg.w + bb.w;
        if (out) *reinterpret_cast<float4*>(out + m * C + c4 * 4) = o;
        if (out_hi) store_split4(out_hi, out_lo, m * C + c4 * 4, o);
      }
    }
  }
}

// ws: [gridDim.x][2][C] per-CTA partial (dg, db)
__global__ void __launch_bounds__(256) ln_backward_kernel(const float* __restrict__ d_out,
                                                          const float* __restrict__ x,
                                                          const float* __restrict__ g,
                                                          const float* __restrict__ d_res,
                                                          float* __restrict__ dx, float* __restrict__ ws,
                                                          int64_t M, int C) {
  extern __shared__ float sred[];   // [8 warps][2][C]
  pdl_wait();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * 8 + wib;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  const int nv = C >> 2;
  float4 adg[LN_MAX_V], adb[LN_MAX_V];
#pragma unroll
  for (int j = 0; j < LN_MAX_V; ++j) {
    adg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    adb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t m = warp; m < M; m += nwarps) {
    float4 v[LN_MAX_V], d[LN_MAX_V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_V; ++j) {
      const int c4 = lane + j * 32;
      if (c4 < nv) {
        v[j] = __ldg(reinterpret_cast<const float4*>(x + m * C + c4 * 4));
        d[j] = __ldg(reinterpret_cast<const float4*>(d_out + m * C + c4 * 4));
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_V; ++j) {
      const int c4 = lane + j * 32;
      if (c4 < nv) {
        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
        q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
      }
    }
    const float stdv = sqrtf(warp_sum(q) / C);
    const float sdn = stdv + kLnEps;
    const float inv = 1.f / sdn;
    // dn = d_out * g ; accumulate param grads ; sums for dx
    float sdnv = 0.f, sdnx = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAX_V; ++j) {
      const int c4 = lane + j * 32;
      if (c4 < nv) {
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g + c4 * 4));
        adg[j].x += d[j].x * v[j].x * inv; adg[j].y += d[j].y * v[j].y * inv;
        adg[j].z += d[j].z * v[j].z * inv; adg[j].w += d[j].w * v[j].w * inv;
        adb[j].x += d[j].x; adb[j].y += d[j].y; adb[j].z += d[j].z; adb[j].w += d[j].w;
        d[j].x *= gg.x; d[j].y *= gg.y; d[j].z *= gg.z; d[j].w *= gg.w;
        sdnv += (d[j].x + d[j].y) + (d[j].z + d[j].w);
        sdnx += (d[j].x * v[j].x + d[j].y * v[j].y) + (d[j].z * v[j].z + d[j].w * v[j].w);
      }
    }
    sdnv = warp_sum(sdnv);
    sdnx = warp_sum(sdnx);
    const float k1 = sdnv / (C * sdn);
    const float k2 = sdnx / ((float)C * fmaxf(stdv, 1e-30f) * sdn * sdn);
#pragma unroll
    for (int j = 0; j < LN_MAX_V; ++j) {
      const int c4 = lane + j * 32;
      if (c4 < nv) {
        float4 o;
        o.x = d[j].x * inv - k1 - v[j].x * k2;
        o.y = d[j].y * inv - k1 - v[j].y * k2;
        o.z = d[j].z * inv - k1 - v[j].z * k2;
        o.w = d[j].w * inv - k1 - v[j].w * k2;
        if (d_res) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(d_res + m * C + c4 * 4));
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        *reinterpret_cast<float4*>(dx + m * C + c4 * 4) = o;
      }
    }
  }
  // CTA-level reduce of parameter partials over the 8 warps (fixed order)
#pragma unroll
  for (int j = 0; j < LN_MAX_V; ++j) {
    const int c4 = lane + j * 32;
    if (c4 < nv) {
      *reinterpret_cast<float4*>(&sred[(wib * 2 + 0) * C + c4 * 4]) = adg[j];
      *reinterpret_cast<float4*>(&sred[(wib * 2 + 1) * C + c4 * 4]) = adb[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int which = i / C, c = i - which * C;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sred[(w * 2 + which) * C + c];
    ws[((int64_t)blockIdx.x * 2 + which) * C + c] = t;
  }
}

// grid (C/32, splits), block (32, 32): 32 row lanes stride over this split's share of the per-CTA partials
__global__ void __launch_bounds__(1024) ln_param_finalize_kernel(const float* __restrict__ ws, int nparts, int C,
                                                                 float* __restrict__ dg, float* __restrict__ db) {
  __shared__ float red[32][2][33];
  pdl_wait();
  const int c = blockIdx.x * 32 + threadIdx.x;
  float tg = 0.f, tb = 0.f;
  if (c < C) {
    for (int p = blockIdx.y * 32 + threadIdx.y; p < nparts; p += 32 * gridDim.y) {
      tg += ws[((int64_t)p * 2 + 0) * C + c];
      tb += ws[((int64_t)p * 2 + 1) * C + c];
    }
  }
  red[threadIdx.y][0][threadIdx.x] = tg;
  red[threadIdx.y][1][threadIdx.x] = tb;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      sg += red[i][0][threadIdx.x];
      sb += red[i][1][threadIdx.x];
    }
    atomicAdd(dg + c, sg);
    atomicAdd(db + c, sb);
  }
}


// ---- specialised LayerNorm kernels: LPP lanes per pixel, V float4 per lane (C = 4 * LPP * V) --------
// The generic kernels above keep LN_MAX_V float4 slots per lane alive (168 registers in backward) and idle
// half a warp at C = 64; these keep exactly what the channel count needs and pack 32/LPP pixels per warp.
template <int LPP>
__device__ __forceinline__ float sub_sum(float v) {
#pragma unroll
  for (int o = LPP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPP, int V>
__global__ void __launch_bounds__(256) ln_forward_t_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                           const float* __restrict__ bta, float* __restrict__ out,
                                                           int64_t M, __nv_bfloat16* __restrict__ out_hi,
                                                           __nv_bfloat16* __restrict__ out_lo) {
  constexpr int C = 4 * LPP * V, PPW = 32 / LPP;
  pdl_wait();
  const int lane = threadIdx.x & 31, sub = lane / LPP, l = lane % LPP;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 gg[V], bb[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    gg[j] = __ldg(reinterpret_cast<const float4*>(g + (l + j * LPP) * 4));
    bb[j] = __ldg(reinterpret_cast<const float4*>(bta + (l + j * LPP) * 4));
  }
  for (int64_t m0 = warp * PPW; m0 < M; m0 += nwarps * PPW) {
    const int64_t m = m0 + sub;
    const bool ok = m < M;
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = ok ? __ldg(reinterpret_cast<const float4*>(x + m * C + (l + j * LPP) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = sub_sum<LPP>(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    const float stdv = sqrtf(sub_sum<LPP>(q) * (1.f / C));
    const float inv = 1.f / (stdv + kLnEps);
    if (ok) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        float4 o;
        o.x = v[j].x * inv * gg[j].x + bb[j].x;
        o.y = v[j].y * inv * gg[j].y + bb[j].y;
        o.z = v[j].z * inv * gg[j].z + bb[j].z;
        o.w = v[j].w * inv * gg[j].w + bb[j].w;
        const int64_t off = m * C + (l + j * LPP) * 4;
        if (out) *reinterpret_cast<float4*>(out + off) = o;
        if (out_hi) store_split4(out_hi, out_lo, off, o);
      }
    }
  }
}

template <int LPP, int V>
__global__ void __launch_bounds__(256) ln_backward_t_kernel(const float* __restrict__ d_out, const float* __restrict__ x,
                                                            const float* __restrict__ g, const float* __restrict__ d_res,
                                                            float* __restrict__ dx, float* __restrict__ ws, int64_t M) {
  constexpr int C = 4 * LPP * V, PPW = 32 / LPP;
  extern __shared__ float sred[];   // [8 warps][2][C]
  pdl_wait();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, sub = lane / LPP, l = lane % LPP;
  const int64_t warp = (int64_t)blockIdx.x * 8 + wib;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  float4 gg[V], adg[V], adb[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    gg[j] = __ldg(reinterpret_cast<const float4*>(g + (l + j * LPP) * 4));
    adg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    adb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t m0 = warp * PPW; m0 < M; m0 += nwarps * PPW) {
    const int64_t m = m0 + sub;
    const bool ok = m < M;
    float4 v[V], d[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int64_t off = m * C + (l + j * LPP) * 4;
      v[j] = ok ? __ldg(reinterpret_cast<const float4*>(x + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
      d[j] = ok ? __ldg(reinterpret_cast<const float4*>(d_out + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = sub_sum<LPP>(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    const float stdv = sqrtf(sub_sum<LPP>(q) * (1.f / C));
    const float sdn = stdv + kLnEps;
    const float inv = 1.f / sdn;
    float sdnv = 0.f, sdnx = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      adg[j].x += d[j].x * v[j].x * inv; adg[j].y += d[j].y * v[j].y * inv;
      adg[j].z += d[j].z * v[j].z * inv; adg[j].w += d[j].w * v[j].w * inv;
      adb[j].x += d[j].x; adb[j].y += d[j].y; adb[j].z += d[j].z; adb[j].w += d[j].w;
      d[j].x *= gg[j].x; d[j].y *= gg[j].y; d[j].z *= gg[j].z; d[j].w *= gg[j].w;
      sdnv += (d[j].x + d[j].y) + (d[j].z + d[j].w);
      sdnx += (d[j].x * v[j].x + d[j].y * v[j].y) + (d[j].z * v[j].z + d[j].w * v[j].w);
    }
    sdnv = sub_sum<LPP>(sdnv);
    sdnx = sub_sum<LPP>(sdnx);
    const float k1 = sdnv / (C * sdn);
    const float k2 = sdnx / ((float)C * fmaxf(stdv, 1e-30f) * sdn * sdn);
    if (ok) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const int64_t off = m * C + (l + j * LPP) * 4;
        float4 o;
        o.x = d[j].x * inv - k1 - v[j].x * k2;
        o.y = d[j].y * inv - k1 - v[j].y * k2;
        o.z = d[j].z * inv - k1 - v[j].z * k2;
        o.w = d[j].w * inv - k1 - v[j].w * k2;
        if (d_res) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(d_res + off));
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        *reinterpret_cast<float4*>(dx + off) = o;
      }
    }
  }
  // fold the PPW pixel sub-groups of the warp (same channels), then the 8 warps, in a fixed order
#pragma unroll
  for (int j = 0; j < V; ++j) {
#pragma unroll
    for (int o = LPP; o < 32; o <<= 1) {
      adg[j].x += __shfl_xor_sync(0xffffffffu, adg[j].x, o); adg[j].y += __shfl_xor_sync(0xffffffffu, adg[j].y, o);
      adg[j].z += __shfl_xor_sync(0xffffffffu, adg[j].z, o); adg[j].w += __shfl_xor_sync(0xffffffffu, adg[j].w, o);
      adb[j].x += __shfl_xor_sync(0xffffffffu, adb[j].x, o); adb[j].y += __shfl_xor_sync(0xffffffffu, adb[j].y, o);
      adb[j].z += __shfl_xor_sync(0xffffffffu, adb[j].z, o); adb[j].w += __shfl_xor_sync(0xffffffffu, adb[j].w, o);
    }
    if (sub == 0) {
      *reinterpret_cast<float4*>(&sred[(wib * 2 + 0) * C + (l + j * LPP) * 4]) = adg[j];
      *reinterpret_cast<float4*>(&sred[(wib * 2 + 1) * C + (l + j * LPP) * 4]) = adb[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int which = i / C, c = i - which * C;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sred[(w * 2 + which) * C + c];
    ws[((int64_t)blockIdx.x * 2 + which) * C + c] = t;
  }
}

}  // namespace

static int check_gn_shape(const LaunchCtx& lc, int C) {
  if (C % 32 != 0 || C > 1024 || C < 32) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "GroupNorm: C must be a multiple of 32 in [32, 1024]");
  return IGM_OK;
}

// pixels per CTA: enough CTAs for a few waves of 8 resident CTAs per SM, never below kGnChunkMin
// (the workspaces are sized for kGnChunkMin) and always a multiple of the pixels in flight per iteration
int gn_chunk(int B, int HW, int C) {
  int chunk = kGnChunk;
  while (chunk > kGnChunkMin && (int64_t)B * cdiv(HW, chunk) < 4096) chunk >>= 1;
  const int ppi = 256 / (C >> 2);
  if (chunk < ppi) chunk = ppi;
  return chunk;
}

int launch_gn_partial(const LaunchCtx& lc, const float* y, int B, int HW, int C, float* part) {
  IGM_TRY(check_gn_shape(lc, C));
  const int kGnChunk = igm::kGnChunk;   // the partial layout of this entry point is fixed (64-pixel chunks)
  const int nchunks = cdiv(HW, kGnChunk);
  ProfScope ps_(lc, K_NORM, 3.0 * B * HW * C, 4.0 * B * HW * C);
  gn_partial_kernel<<<B * nchunks, 256, 0, lc.stream>>>(y, HW, C, nchunks, part, kGnChunk);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_gn_apply(const LaunchCtx& lc, const float* y, const float* part, const float* gamma,
                    const float* beta, const float* temb, int temb_stride, const float* res, float* out,
                    float* stats, int B, int HW, int C, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int nparts) {
  IGM_TRY(check_gn_shape(lc, C));
  if (nparts <= 0) nparts = cdiv(HW, igm::kGnChunk);   // partials written by launch_gn_partial
  const int L = C >> 2;
  static const bool fast_off = [] { const char* e = getenv("IGM_GN_FAST"); return e && e[0] == '0'; }();
  if (!fast_off && (L & (L - 1)) == 0 && L >= 8 && L <= 256) {
    int lgL = 3;
    while ((1 << lgL) < L) ++lgL;
    const int chunk = (256 >> lgL) * kGnIter;
    ProfScope ps_(lc, K_NORM, 30.0 * B * HW * C, 4.0 * B * HW * C * (res ? 3 : 2));
    cudaError_t le = launch_pdl(gn_apply_fast_kernel, dim3(cdiv(HW, chunk), B), dim3(256), 0, lc.stream, y, part, gamma, beta, temb,
                                temb_stride, res, out, stats, HW, C, lgL, out_hi, out_lo, nparts,
                                1.0 / ((double)HW * (C / kGroups)));
    if (le != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le));
    IGM_POST_LAUNCH(lc);
    return IGM_OK;
  }
  const int kGnChunk = gn_chunk(B, HW, C);
  const int nchunks = cdiv(HW, kGnChunk);
  ProfScope ps_(lc, K_NORM, 30.0 * B * HW * C, 4.0 * B * HW * C * (res ? 3 : 2));
  cudaError_t le = launch_pdl(gn_apply_kernel, dim3(B * nchunks), dim3(256), 0, lc.stream, y, part, gamma, beta, temb, temb_stride,
                              res, out, stats, HW, C, nchunks, out_hi, out_lo, nparts, kGnChunk);
  if (le != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le));
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

// cluster size and float4-per-thread of the fused backward (0: shape not covered -> two-pass path)
static int gn_fused_cluster(const GnBwdArgs& a, int& nv) {
  const int64_t E = (int64_t)a.HW * a.C;
  if (a.dtemb && (a.dtemb_stride % 4 != 0)) return 0;
  const int ppi = 256 / (a.C >> 2);
  for (int v : {8, 4, 2, 1}) {   // large per-thread tiles first: clusters of 8 measured slower than 8 float4 per thread
    if (E % (1024 * v) != 0) continue;
    const int64_t cs = E / (1024 * v);
    if (cs != 1 && cs != 2 && cs != 4 && cs != 8) continue;
    if (a.HW % cs != 0 || a.HW / cs != v * ppi) continue;
    nv = v;
    return (int)cs;
  }
  return 0;
}

static bool gn_fused_disabled() {
  static const bool off = [] { const char* e = getenv("IGM_GN_FUSED"); return e && e[0] == '0'; }();
  return off;
}

bool gn_backward_is_fused(const GnBwdArgs& a) {
  int nv = 0;
  return !gn_fused_disabled() && a.C % 32 == 0 && a.C >= 32 && a.C <= 1024 && gn_fused_cluster(a, nv) > 0;
}

int launch_gn_backward(const LaunchCtx& lc, const GnBwdArgs& a) {
  IGM_TRY(check_gn_shape(lc, a.C));
  const bool fused_off = gn_fused_disabled();
  int nv = 0;
  const int cs = fused_off ? 0 : gn_fused_cluster(a, nv);
  if (cs == 0 && (a.dout_hi || a.dout_colsum)) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "GroupNorm backward: d_out by-products need the fused kernel");
  if (cs > 0) {
    ProfScope ps_(lc, K_NORM, 80.0 * a.B * a.HW * a.C, 4.0 * a.B * a.HW * a.C * 3);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(a.B * cs));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = lc.stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t e;
    static const bool bulk = [] { const char* v = getenv("IGM_GN_BULK"); return !(v && v[0] == '0'); }();
    // bulk copies need 16-byte aligned sources: every tensor comes from the 256-byte aligned arena, slices are 4 KB multiples
    const bool aligned = ((reinterpret_cast<uintptr_t>(a.y) | reinterpret_cast<uintptr_t>(a.d_out)) & 15) == 0;
    if (bulk && aligned) {
      const int smem = std::max(2 * nv * 1024 * (int)sizeof(float), 256 * 4 * (int)sizeof(float4));
      static bool attr_set = false;
      if (!attr_set) {
        e = cudaFuncSetAttribute(gn_bwd_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
        attr_set = true;
      }
      cfg.dynamicSmemBytes = (size_t)smem;
      e = cudaLaunchKernelEx(&cfg, gn_bwd_bulk_kernel, a, cs, nv);
      if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
      IGM_POST_LAUNCH(lc);
      return IGM_OK;
    }
    switch (nv) {
      case 1: e = cudaLaunchKernelEx(&cfg, gn_bwd_fused_kernel<1>, a, cs); break;
      case 2: e = cudaLaunchKernelEx(&cfg, gn_bwd_fused_kernel<2>, a, cs); break;
      case 4: e = cudaLaunchKernelEx(&cfg, gn_bwd_fused_kernel<4>, a, cs); break;
      default: e = cudaLaunchKernelEx(&cfg, gn_bwd_fused_kernel<8>, a, cs); break;
    }
    if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
    IGM_POST_LAUNCH(lc);
    return IGM_OK;
  }
  const int kGnChunk = gn_chunk(a.B, a.HW, a.C);
  const int nchunks = cdiv(a.HW, kGnChunk);
  ProfScope ps_(lc, K_NORM, 80.0 * a.B * a.HW * a.C, 4.0 * a.B * a.HW * a.C * 5);
  gn_bwd_reduce_kernel<<<a.B * nchunks, 256, 0, lc.stream>>>(a, nchunks, kGnChunk);
  IGM_POST_LAUNCH(lc);
  gn_bwd_apply_kernel<<<a.B * nchunks, 256, 0, lc.stream>>>(a, nchunks, kGnChunk);
  IGM_POST_LAUNCH(lc);
  gn_bwd_param_kernel<<<dim3(cdiv(a.C, 32), a.B), dim3(32, 8), 0, lc.stream>>>(a, nchunks);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

static int ln_grid(int64_t M) {
  int64_t g = cdiv64(M, 8 * 4);
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

int launch_ln_forward(const LaunchCtx& lc, const float* x, const float* g, const float* b, float* out,
                      int64_t M, int C, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo) {
  if (C % 4 != 0 || C > 1024) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "LayerNorm: C must be a multiple of 4, <= 1024");
  ProfScope ps_(lc, K_NORM, 8.0 * M * C, 8.0 * M * C);
  const int grid = ln_grid(M);
  switch (C) {
    case 32:   { cudaError_t le_ = launch_pdl(ln_forward_t_kernel<8, 1>, dim3(grid), dim3(256), (size_t)(0), lc.stream, x, g, b, out, M, out_hi, out_lo); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
    case 64:   { cudaError_t le_ = launch_pdl(ln_forward_t_kernel<16, 1>, dim3(grid), dim3(256), (size_t)(0), lc.stream, x, g, b, out, M, out_hi, out_lo); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
    case 128:  { cudaError_t le_ = launch_pdl(ln_forward_t_kernel<32, 1>, dim3(grid), dim3(256), (size_t)(0), lc.stream, x, g, b, out, M, out_hi, out_lo); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
    case 256:  { cudaError_t le_ = launch_pdl(ln_forward_t_kernel<32, 2>, dim3(grid), dim3(256), (size_t)(0), lc.stream, x, g, b, out, M, out_hi, out_lo); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
    case 512:  { cudaError_t le_ = launch_pdl(ln_forward_t_kernel<32, 4>, dim3(grid), dim3(256), (size_t)(0), lc.stream, x, g, b, out, M, out_hi, out_lo); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
    default:   { cudaError_t le_ = launch_pdl(ln_forward_kernel, dim3(grid), dim3(256), (size_t)(0), lc.stream, x, g, b, out, M, C, out_hi, out_lo); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
  }
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int ln_backward_parts(int64_t M) { return ln_grid(M); }

// parameter gradients from the per-CTA partials of launch_ln_backward (split off so that it can leave the critical path)
int launch_ln_param_finalize(const LaunchCtx& lc, const float* ws, int64_t M, int C, float* dg, float* db) {
  const int grid = ln_grid(M);
  ProfScope ps_(lc, K_NORM, 2.0 * grid * C, 8.0 * grid * C);
  { cudaError_t le_ = launch_pdl(ln_param_finalize_kernel, dim3(cdiv(C, 32), cdiv(grid, 128)), dim3(32, 32), (size_t)0, lc.stream, ws, grid, C, dg, db); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); }
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_ln_backward(const LaunchCtx& lc, const float* d_out, const float* x, const float* g,
                       const float* d_res, float* dx, float* dg, float* db, float* ws, int64_t M, int C, bool finalize) {
  if (C % 4 != 0 || C > 1024) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "LayerNorm: C must be a multiple of 4, <= 1024");
  const int grid = ln_grid(M);
  ProfScope ps_(lc, K_NORM, 20.0 * M * C, 4.0 * M * C * (d_res ? 4 : 3));
  const size_t smem = (size_t)8 * 2 * C * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ln_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
  }
  switch (C) {
    case 32:   { cudaError_t le_ = launch_pdl(ln_backward_t_kernel<8, 1>, dim3(grid), dim3(256), (size_t)(smem), lc.stream, d_out, x, g, d_res, dx, ws, M); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
    case 64:   { cudaError_t le_ = launch_pdl(ln_backward_t_kernel<16, 1>, dim3(grid), dim3(256), (size_t)(smem), lc.stream, d_out, x, g, d_res, dx, ws, M); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
    case 128:  { cudaError_t le_ = launch_pdl(ln_backward_t_kernel<32, 1>, dim3(grid), dim3(256), (size_t)(smem), lc.stream, d_out, x, g, d_res, dx, ws, M); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
    case 256:  { cudaError_t le_ = launch_pdl(ln_backward_t_kernel<32, 2>, dim3(grid), dim3(256), (size_t)(smem), lc.stream, d_out, x, g, d_res, dx, ws, M); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
    case 512:  { cudaError_t le_ = launch_pdl(ln_backward_t_kernel<32, 4>, dim3(grid), dim3(256), (size_t)(smem), lc.stream, d_out, x, g, d_res, dx, ws, M); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
    default:   { cudaError_t le_ = launch_pdl(ln_backward_kernel, dim3(grid), dim3(256), (size_t)(smem), lc.stream, d_out, x, g, d_res, dx, ws, M, C); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); } break;
  }
  IGM_POST_LAUNCH(lc);
  if (!finalize) return IGM_OK;
  { cudaError_t le_ = launch_pdl(ln_param_finalize_kernel, dim3(cdiv(C, 32), cdiv(grid, 128)), dim3(32, 32), (size_t)0, lc.stream, ws, grid, C, dg, db); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); }
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace igm
