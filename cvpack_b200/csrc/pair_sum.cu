// Group-pair sums with analytic gradients: NumberOfContacts, AttractionStrength, ShortestDistance.
//
// Replaces, for a whole batch of frames, what OpenMM's CustomNonbondedForce with one interaction
// group does per frame for cvpack/number_of_contacts.py:143-161, attraction_strength.py:189-231 and
// shortest_distance.py:122-138.
//
// Pipeline per chunk of frames (all on one stream, deterministic -- no floating-point atomics):
//   1. pack      gather group coordinates into contiguous 16/32-byte records [B][n] (x,y,z,tag)
//   2. sweep A   i = group1 in registers, j = group2 streamed through shared memory by TMA bulk
//                copies (cp.async.bulk + mbarrier, double buffered): energy partials + i gradients
//   3. sweep B   roles swapped: gradients of group2 atoms (skipped for value-only calls)
//   4. finish    deterministic reduction of the per-block energy partials, exclusions, scaling
//   5. scatter   packed gradients -> [B][N][3]
#include "pair_common.cuh"
#include "pair_cells.cuh"

namespace cvp {

namespace {

constexpr int PAIR_THREADS = 128;
constexpr int PAIR_IPT = 2;      // i atoms per thread
constexpr int PAIR_TILE = 256;   // j atoms per shared-memory stage
constexpr int PAIR_I_PER_BLOCK = PAIR_THREADS * PAIR_IPT;

// ---- 1. pack -----------------------------------------------------------------------------------
template <typename T>
__global__ void pack_groups_kernel(const T* __restrict__ pos, int64_t num_atoms,
                                   const int32_t* __restrict__ idx1, const int32_t* __restrict__ tag1,
                                   int n1, const int32_t* __restrict__ idx2,
                                   const int32_t* __restrict__ tag2, int n2,
                                   typename Vec4<T>::type* __restrict__ out1,
                                   typename Vec4<T>::type* __restrict__ out2, int64_t total) {
  using T4 = typename Vec4<T>::type;
  const int n = n1 + n2;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / n;
    int k = (int)(t - b * n);
    bool first = k < n1;
    int slot = first ? k : k - n1;
    int atom = first ? idx1[slot] : idx2[slot];
    int tag = first ? tag1[slot] : tag2[slot];
    const T* p = pos + (b * num_atoms + atom) * 3;
    T4 rec;
    rec.x = p[0];
    rec.y = p[1];
    rec.z = p[2];
    rec.w = tag_encode(tag, T(0));
    if (first)
      out1[b * n1 + slot] = rec;
    else
      out2[b * n2 + slot] = rec;
  }
}

// ---- 2/3. sweep --------------------------------------------------------------------------------
template <typename T, int KIND, int PBC, bool OVERLAP, bool GRAD, bool FAST6 = false>
__global__ void __launch_bounds__(PAIR_THREADS)
    pair_sweep_kernel(const typename Vec4<T>::type* __restrict__ ipos,
                      const typename Vec4<T>::type* __restrict__ jpos,
                      const typename Vec4<T>::type* __restrict__ iprm,
                      const typename Vec4<T>::type* __restrict__ jprm, int ni, int nj, int nblk_i,
                      const T* __restrict__ box, int box_mode, PairConsts<T> pc,
                      const typename PairFn<T, KIND>::Acc* __restrict__ frame_coef,
                      double* __restrict__ energy_partials,
                      typename Vec4<T>::type* __restrict__ igrad) {
  using T4 = typename Vec4<T>::type;
  using Fn = PairFn<T, KIND>;
  using Acc = typename Fn::Acc;
  constexpr int NCH = Fn::NCH;
  constexpr bool PRM = Fn::HAS_PARAMS;

  __shared__ __align__(128) T4 jbuf[2][PAIR_TILE];
  __shared__ __align__(128) T4 pbuf[PRM ? 2 : 1][PRM ? PAIR_TILE : 1];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ double scratch[NCH * (PAIR_THREADS / 32)];

  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x / nblk_i;
  const int ib = blockIdx.x - (int)b * nblk_i;

  Box<T> bx;
  if (PBC != 0) bx = load_box(box, box_mode, b);

  const T4* jp = jpos + b * nj;
  const int ntile = (nj + PAIR_TILE - 1) / PAIR_TILE;

  auto issue = [&](int tile) {
    const int s = tile & 1;
    const int count = min(PAIR_TILE, nj - tile * PAIR_TILE);
    const uint32_t bytes = (uint32_t)(count * sizeof(T4));
    mbar_expect_tx(&bars[s], PRM ? 2 * bytes : bytes);
    tma_load_1d(&jbuf[s][0], jp + (size_t)tile * PAIR_TILE, bytes, &bars[s]);
    if (PRM) tma_load_1d(&pbuf[s][0], jprm + (size_t)tile * PAIR_TILE, bytes, &bars[s]);
  };

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    issue(0);
    if (ntile > 1) issue(1);
  }

  // my i atoms
  T xi[PAIR_IPT], yi[PAIR_IPT], zi[PAIR_IPT];
  int tagi[PAIR_IPT];
  T4 prmi[PAIR_IPT];
  bool valid[PAIR_IPT];
  T gx[PAIR_IPT], gy[PAIR_IPT], gz[PAIR_IPT];
  Acc en[PAIR_IPT][NCH];
#pragma unroll
  for (int k = 0; k < PAIR_IPT; ++k) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) en[k][c] = Acc(0);
    int i = ib * PAIR_I_PER_BLOCK + k * PAIR_THREADS + tid;
    valid[k] = i < ni;
    if (!valid[k]) i = ni - 1;
    T4 rec = ipos[b * ni + i];
    xi[k] = rec.x;
    yi[k] = rec.y;
    zi[k] = rec.z;
    tagi[k] = tag_decode(rec.w);
    if (PRM) prmi[k] = iprm[i];
    gx[k] = gy[k] = gz[k] = T(0);
  }
  const Acc* fc = (KIND == CVP_PAIR_SOFTMIN && GRAD) ? frame_coef + 2 * b : nullptr;

  for (int tile = 0; tile < ntile; ++tile) {
    const int s = tile & 1;
    mbar_wait(&bars[s], (tile >> 1) & 1);
    const int count = min(PAIR_TILE, nj - tile * PAIR_TILE);
#pragma unroll 4
    for (int jj = 0; jj < count; ++jj) {
      const T4 pj = jbuf[s][jj];
      T4 qj;
      if (PRM) qj = pbuf[s][jj];
      const int tagj = OVERLAP ? tag_decode(pj.w) : 0;
#pragma unroll
      for (int k = 0; k < PAIR_IPT; ++k) {
        T dx = pj.x - xi[k], dy = pj.y - yi[k], dz = pj.z - zi[k];
        pair_min_image<PBC>(dx, dy, dz, bx);
        T r2 = dx * dx + dy * dy + dz * dz;
        if (r2 < pc.cutoff2) {
          T w = T(1);
          if (OVERLAP) {
            if (((tagi[k] ^ tagj) & TAG_INDEX_MASK) == 0) continue;  // same atom
            if (tagi[k] & tagj & TAG_BOTH) w = T(0.5);               // pair seen from both sides
          }
          Acc e[NCH];
          T coef;
          if constexpr (KIND == CVP_PAIR_SOFTMIN) {
            // displacement in double from the stored coordinates (see PairFn<SOFTMIN>)
            double ex = (double)pj.x - (double)xi[k], ey = (double)pj.y - (double)yi[k],
                   ez = (double)pj.z - (double)zi[k];
            if (PBC != 0) {
              Box<double> bd;
              bd.ax = bx.ax; bd.bx = bx.bx; bd.by = bx.by; bd.cx = bx.cx; bd.cy = bx.cy; bd.cz = bx.cz;
              bd.inv_ax = 1.0 / bd.ax; bd.inv_by = 1.0 / bd.by; bd.inv_cz = 1.0 / bd.cz;
              min_image<PBC>(ex, ey, ez, bd);
            }
            Fn::eval_d(pc, ex, ey, ez, fc, e, coef);
          } else if constexpr (FAST6) {
            contacts_rational6(pc, r2, e[0], coef);  // 1/(1+x^6): r^2 only, no square root
          } else
            Fn::eval(pc, r2, prmi[k], qj, fc, e, coef);
#pragma unroll
          for (int c = 0; c < NCH; ++c) en[k][c] += Acc(w) * e[c];
          if (GRAD) {
            coef *= w;
            gx[k] -= coef * dx;
            gy[k] -= coef * dy;
            gz[k] -= coef * dz;
          }
        }
      }
    }
    __syncthreads();  // everyone is done with stage s
    if (tid == 0 && tile + 2 < ntile) issue(tile + 2);
  }

  if (GRAD) {
#pragma unroll
    for (int k = 0; k < PAIR_IPT; ++k) {
      int i = ib * PAIR_I_PER_BLOCK + k * PAIR_THREADS + tid;
      if (valid[k]) {
        T4 g;
        g.x = gx[k];
        g.y = gy[k];
        g.z = gz[k];
        g.w = T(0);
        igrad[b * ni + i] = g;
      }
    }
  }
  if (energy_partials != nullptr) {
    // atoms clamped to the last valid slot (tail of the last block) must not contribute
    double v[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      v[c] = 0.0;
#pragma unroll
      for (int k = 0; k < PAIR_IPT; ++k) v[c] += valid[k] ? (double)en[k][c] : 0.0;
    }
    block_sum<NCH>(v, scratch);
    if (tid == 0) {
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        energy_partials[((size_t)b * nblk_i + ib) * NCH + c] = v[c];
    }
  }
}

// ---- exclusions (pairs never counted): computed directly from the frame, subtracted later -------
// One thread per (frame, listed atom): walks the partners of that atom in CSR order, so the
// gradient correction has a single writer per element and a fixed summation order.
template <typename T, int KIND, int PBC>
__global__ void exclusion_gradient_kernel(const T* __restrict__ pos, int64_t num_atoms,
                                          const typename Vec4<T>::type* __restrict__ atom_params,
                                          const int32_t* __restrict__ ex_atoms,
                                          const int32_t* __restrict__ ex_offsets,
                                          const int32_t* __restrict__ ex_partners, int num_ex_atoms,
                                          const T* __restrict__ box, int box_mode, PairConsts<T> pc,
                                          T scale, T* __restrict__ grad, int64_t total) {
  using T4 = typename Vec4<T>::type;
  using Fn = PairFn<T, KIND>;
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  int64_t b = t / num_ex_atoms;
  int k = (int)(t - b * num_ex_atoms);
  int a = ex_atoms[k];
  Box<T> bx;
  if (PBC != 0) bx = load_box(box, box_mode, b);
  const T* pa = pos + (b * num_atoms + a) * 3;
  T4 qa;
  if (Fn::HAS_PARAMS) qa = atom_params[a];
  T gx = 0, gy = 0, gz = 0;
  for (int e = ex_offsets[k]; e < ex_offsets[k + 1]; ++e) {
    int p = ex_partners[e];
    const T* pp = pos + (b * num_atoms + p) * 3;
    T dx = pp[0] - pa[0], dy = pp[1] - pa[1], dz = pp[2] - pa[2];
    pair_min_image<PBC>(dx, dy, dz, bx);
    T r2 = dx * dx + dy * dy + dz * dz;
    if (r2 < pc.cutoff2) {
      T4 qp;
      if (Fn::HAS_PARAMS) qp = atom_params[p];
      typename Fn::Acc en[Fn::NCH];
      T coef;
      Fn::eval(pc, r2, qa, qp, nullptr, en, coef);
      gx -= coef * dx;
      gy -= coef * dy;
      gz -= coef * dz;
    }
  }
  T* g = grad + (b * num_atoms + a) * 3;
  g[0] -= scale * gx;
  g[1] -= scale * gy;
  g[2] -= scale * gz;
}

// One block per frame: energy of all excluded pairs (each unordered pair listed once).
template <typename T, int KIND, int PBC>
__global__ void __launch_bounds__(128)
    exclusion_energy_kernel(const T* __restrict__ pos, int64_t num_atoms,
                            const typename Vec4<T>::type* __restrict__ atom_params,
                            const int32_t* __restrict__ pairs, int num_pairs,
                            const T* __restrict__ box, int box_mode, PairConsts<T> pc,
                            double* __restrict__ out) {
  using T4 = typename Vec4<T>::type;
  using Fn = PairFn<T, KIND>;
  __shared__ double scratch[4];
  const int64_t b = blockIdx.x;
  Box<T> bx;
  if (PBC != 0) bx = load_box(box, box_mode, b);
  double v[1] = {0.0};
  for (int e = threadIdx.x; e < num_pairs; e += blockDim.x) {
    int a = pairs[2 * e], p = pairs[2 * e + 1];
    const T* pa = pos + (b * num_atoms + a) * 3;
    const T* pp = pos + (b * num_atoms + p) * 3;
    T dx = pp[0] - pa[0], dy = pp[1] - pa[1], dz = pp[2] - pa[2];
    pair_min_image<PBC>(dx, dy, dz, bx);
    T r2 = dx * dx + dy * dy + dz * dz;
    if (r2 < pc.cutoff2) {
      T4 qa, qp;
      if (Fn::HAS_PARAMS) {
        qa = atom_params[a];
        qp = atom_params[p];
      }
      typename Fn::Acc en[Fn::NCH];
      T coef;
      Fn::eval(pc, r2, qa, qp, nullptr, en, coef);
      v[0] += (double)en[0];
    }
  }
  block_sum<1>(v, scratch);
  if (threadIdx.x == 0) out[b] = v[0];
}

// ---- 4. finish ---------------------------------------------------------------------------------
// One warp per frame: fixed-order sum of the per-block partials, then the CV value.
template <typename T, int KIND>
__global__ void pair_finish_kernel(const double* __restrict__ partials, int nblk,
                                   const double* __restrict__ excluded, double scale,
                                   double rc, double exp_neg_beta, T* __restrict__ value,
                                   typename PairFn<T, KIND>::Acc* __restrict__ frame_coef,
                                   int accumulate, int64_t num_frames) {
  constexpr int NCH = PairFn<T, KIND>::NCH;
  const int lane = threadIdx.x & 31;
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (b >= num_frames) return;
  double sum[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    double v = 0.0;
    for (int k = lane; k < nblk; k += 32) v += partials[((size_t)b * nblk + k) * NCH + c];
    sum[c] = warp_sum(v);
  }
  if (lane != 0) return;
  double result;
  if (KIND == CVP_PAIR_SOFTMIN) {
    double denom = exp_neg_beta + sum[NCH - 1];
    result = (rc * exp_neg_beta + sum[0]) / denom;
    if (frame_coef != nullptr) {
      using Acc = typename PairFn<T, KIND>::Acc;
      frame_coef[2 * b] = Acc(1.0 / denom);
      frame_coef[2 * b + 1] = Acc(-result / denom);
    }
  } else {
    double e = sum[0];
    if (excluded != nullptr) e -= excluded[b];
    result = scale * e;
  }
  value[b] = accumulate ? T((double)value[b] + result) : T(result);
}

// ---- 5. scatter --------------------------------------------------------------------------------
template <typename T>
__global__ void scatter_add_kernel(const typename Vec4<T>::type* __restrict__ packed,
                                   const int32_t* __restrict__ idx, int n, int64_t num_atoms,
                                   T scale, T* __restrict__ grad, int64_t total) {
  using T4 = typename Vec4<T>::type;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / n;
    int slot = (int)(t - b * n);
    T4 g = packed[t];
    T* out = grad + (b * num_atoms + idx[slot]) * 3;
    out[0] += scale * g.x;
    out[1] += scale * g.y;
    out[2] += scale * g.z;
  }
}

// ---- host side ---------------------------------------------------------------------------------
struct PairSumCV : cvp_cv {
  cvp_pair_sum_desc d{};
  std::vector<int32_t> g1, g2, tag1, tag2;
  std::vector<int32_t> ex_pairs;                       // unique unordered pairs, flattened
  std::vector<int32_t> ex_atoms, ex_offsets, ex_partners;  // CSR by atom
  std::vector<double> charge, sigma, epsilon, sign;
  bool overlap = false;
  bool identical = false;  // group1 and group2 are the same set
  int use_cells = 1;       // option "cells": 1 = cell-sorted fast path when eligible, 0 = all pairs
  int nsplit_override = 0; // option "nsplit": CTAs per frame in the cell sweep (0 = automatic)
  int use_sym = 1;         // option "sym": single symmetric sweep when eligible (pair_cells.cuh)
  DeviceBuffer d_g1, d_g2, d_tag1, d_tag2, d_ex_pairs, d_ex_atoms, d_ex_offsets, d_ex_partners;
  DeviceBuffer d_prm1[2], d_prm2[2], d_prm_atom[2];  // per dtype

  ~PairSumCV() override {
    for (DeviceBuffer* buf : {&d_g1, &d_g2, &d_tag1, &d_tag2, &d_ex_pairs, &d_ex_atoms,
                              &d_ex_offsets, &d_ex_partners, &d_prm1[0], &d_prm1[1], &d_prm2[0],
                              &d_prm2[1], &d_prm_atom[0], &d_prm_atom[1]})
      buf->release();
  }

  int nblk(int n) const { return (int)div_up(n, PAIR_I_PER_BLOCK); }
  double periodic_cutoff() const override { return d.periodic && d.cutoff > 0 ? d.cutoff : 0.0; }

  template <typename T>
  int upload_params(int which) {
    using T4 = typename Vec4<T>::type;
    if (d.kind != CVP_PAIR_ATTRACTION) return CVP_OK;
    auto make = [&](const std::vector<int32_t>* group, DeviceBuffer& buffer) -> int {
      size_t n = group ? group->size() : (size_t)num_atoms;
      std::vector<T4> host(n);
      for (size_t k = 0; k < n; ++k) {
        int a = group ? (*group)[k] : (int)k;
        host[k].x = (T)charge[a];
        host[k].y = (T)sigma[a];
        host[k].z = (T)epsilon[a];
        host[k].w = (T)(sign.empty() ? 1.0 : sign[a]);
      }
      return cvp::upload(buffer, host);
    };
    int status;
    if ((status = make(&g1, d_prm1[which])) != CVP_OK) return status;
    if ((status = make(&g2, d_prm2[which])) != CVP_OK) return status;
    return make(nullptr, d_prm_atom[which]);
  }

  int upload() override {
    int status;
    if ((status = cvp::upload(d_g1, g1)) != CVP_OK) return status;
    if ((status = cvp::upload(d_g2, g2)) != CVP_OK) return status;
    if ((status = cvp::upload(d_tag1, tag1)) != CVP_OK) return status;
    if ((status = cvp::upload(d_tag2, tag2)) != CVP_OK) return status;
    if ((status = cvp::upload(d_ex_pairs, ex_pairs)) != CVP_OK) return status;
    if ((status = cvp::upload(d_ex_atoms, ex_atoms)) != CVP_OK) return status;
    if ((status = cvp::upload(d_ex_offsets, ex_offsets)) != CVP_OK) return status;
    if ((status = cvp::upload(d_ex_partners, ex_partners)) != CVP_OK) return status;
    if ((status = upload_params<float>(0)) != CVP_OK) return status;
    return upload_params<double>(1);
  }

  size_t workspace_per_frame(int dtype, bool need_grad) const override {
    size_t rec = dtype == CVP_F32 ? 16 : 32;
    size_t real = dtype == CVP_F32 ? 4 : 8;
    // sized for the cell path (padded groups + cluster boxes), which also covers the all-pairs path
    // (the smallest chunk any sweep uses gives the most padding)
    const int jc = std::min({cell_jchunk<float>(true), cell_jchunk<float>(false), cell_jchunk<double>(false)});
    size_t n = cell_layout((int)g1.size(), jc).n_pad + cell_layout((int)g2.size(), jc).n_pad;
    size_t bytes = n * rec + 2 * (n / CELL_CLUSTER) * rec + 1024;  // positions + cluster boxes
    if (need_grad) bytes += n * rec + 512;                         // packed gradients
    bytes += (size_t)std::max(nblk((int)g1.size()), 64) * 2 * sizeof(double);  // energy partials
    bytes += sizeof(double) + 2 * sizeof(double) + 0 * real;       // excluded energy, frame coef
    return bytes;
  }
  int last_path = -1;      // statistic "last_path": 0 all pairs, 1 cells (two sweeps), 2 cells (symmetric)
  int get_stat(const char* name, double* out) override {
    if (std::string(name) == "last_path") {
      out[0] = last_path;
      out[1] = 0;
      return CVP_OK;
    }
    return cvp_cv::get_stat(name, out);
  }
  int set_option(const char* name, int value) override {
    const std::string key(name);
    if (key == "cells") {
      std::swap(use_cells, value);
      return value;
    }
    if (key == "nsplit") {
      std::swap(nsplit_override, value);
      return value;
    }
    if (key == "sym") {
      std::swap(use_sym, value);
      return value;
    }
    return cvp_cv::set_option(name, value);
  }
  size_t cell_sort_smem(int n) const {
    return (size_t)(2 * CELL_MAX_CELLS + 1) * sizeof(int) + 2 * (size_t)cell_round_up(n, 8) * sizeof(uint16_t);
  }
  bool cells_eligible(int pbc) const {
    const size_t limit = 200 * 1024;
    return use_cells && d.cutoff > 0 && pbc != 2 && g1.size() <= 65535 && g2.size() <= 65535 &&
           cell_sort_smem((int)std::max(g1.size(), g2.size())) <= limit;
  }
  size_t workspace_fixed(int) const override { return 8 * 256; }
  // run() writes to its workspace only (the tables of the spec are read-only after upload)
  bool concurrent_runs() const override { return true; }
  int64_t max_frames_per_run() const override {
    int64_t blocks = std::max(nblk((int)g1.size()), nblk((int)g2.size()));
    return std::max<int64_t>(1, (int64_t(1) << 31) / blocks - 1);
  }

  template <typename T>
  PairConsts<T> consts() const {
    PairConsts<T> pc{};
    bool has_cutoff = d.cutoff > 0;
    pc.cutoff2 = has_cutoff ? (T)(d.cutoff * d.cutoff) : (T)3.0e38;
    bool sw = d.switch_distance > 0 && has_cutoff && d.switch_distance < d.cutoff;
    pc.has_switch = sw ? 1 : 0;
    pc.rs = sw ? (T)d.switch_distance : (T)0;
    pc.inv_range = sw ? (T)(1.0 / (d.cutoff - d.switch_distance)) : (T)0;
    pc.neg30_inv_range = sw ? (T)(-30.0 / (d.cutoff - d.switch_distance)) : (T)0;
    pc.inv_r0 = d.r0 > 0 ? (T)(1.0 / d.r0) : (T)1;
    pc.inv_r0sq = d.r0 > 0 ? (T)(1.0 / (d.r0 * d.r0)) : (T)1;
    pc.step.kind = d.step_kind;
    pc.step.n = (int)d.step_p0;
    pc.step.m = (int)d.step_p1;
    pc.inv_rc = has_cutoff ? (T)(1.0 / d.cutoff) : (T)1;
    pc.ke = (T)d.coulomb_constant;
    pc.neg_beta_over_rc = has_cutoff ? -d.beta / d.cutoff : -d.beta;
    return pc;
  }

  // Largest |dE/dr| of the pair function 1/(1+x^6) * switch on (0, cutoff), sampled in double.
  double contacts6_max_slope() const {
    const double r0 = d.r0, rc = d.cutoff, rs = d.switch_distance;
    const bool sw_on = rs > 0 && rs < rc;
    double worst = 0.0;
    const int samples = 1 << 14;
    for (int k = 1; k <= samples; ++k) {
      const double r = rc * k / samples;
      const double x = r / r0, x6 = pow(x, 6);
      const double s = 1.0 / (1.0 + x6), ds = -6.0 * x6 / x * s * s / r0;
      double sw = 1.0, dsw = 0.0;
      if (sw_on && r > rs) {
        const double t = (r - rs) / (rc - rs);
        sw = 1.0 - t * t * t * (10.0 - 15.0 * t + 6.0 * t * t);
        dsw = -30.0 * t * t * (1.0 - t) * (1.0 - t) / (rc - rs);
      }
      worst = std::max(worst, std::fabs(ds * sw + s * dsw));
    }
    return worst * 1.01;
  }
  // Decides whether run_cells may use the symmetric sweep and picks the fixed-point scale, a power of
  // two such that one contribution fits comfortably (scale * max|dE/dr| < 2^30) and the wrap counters
  // of pair_cells.cuh cannot leave their 10-bit fields (scale * max|dE/dr| * n_i < 500 * 2^32); the
  // sum itself may wrap, that is what the counters are for.  One pass over the i tiles.
  template <typename T, int KIND, bool OVERLAP>
  bool sym_plan(bool need_grad, int nsplit, int ni_pad, float* scale, int* passes) const {
    if (!std::is_same<T, float>::value || KIND != CVP_PAIR_CONTACTS || OVERLAP) return false;
    if (!use_sym || !need_grad || nsplit != 1) return false;
    if (d.step_kind != CVP_STEP_RATIONAL || (int)d.step_p0 != 6 || !(d.r0 > 0) || !(d.cutoff > 0))
      return false;
    const double slope = contacts6_max_slope();
    if (!(slope > 0)) return false;
    const double single = 1073741824.0 / slope;
    const double limit = std::min({67108864.0, single,
                                   500.0 * 4294967296.0 / (slope * cell_round_up(ni_pad, CELL_TILE))});
    const int exponent = (int)std::floor(std::log2(limit * 0.999));
    if (exponent < 16) return false;
    *scale = (float)std::ldexp(1.0, exponent);
    *passes = 1;
    return true;
  }

  // ---- cell-sorted fast path (pair_cells.cuh) ----
  template <typename T, int KIND, int PBC, bool OVERLAP>
  int run_cells(const EvalArgs& a, void* ws) {
    using T4 = typename Vec4<T>::type;
    constexpr int NCH = PairFn<T, KIND>::NCH;
    constexpr int CPBC = PBC == 2 ? 0 : PBC;  // triclinic never gets here
    const int which = std::is_same<T, float>::value ? 0 : 1;
    const int64_t B = a.num_frames;
    const int64_t N = a.num_atoms;
    const int n1 = (int)g1.size(), n2 = (int)g2.size();
    const bool need_grad = a.grad != nullptr;
    const bool accumulate = (a.flags & CVP_FLAG_ACCUMULATE) != 0;
    const PairConsts<T> pc = consts<T>();
    const T* pos = static_cast<const T*>(a.positions);
    const T* box = static_cast<const T*>(a.box);
    T* value = static_cast<T*>(a.value);
    T* grad = static_cast<T*>(a.grad);
    cudaStream_t st = a.stream;

    int nsplit = nsplit_override;
    if (nsplit <= 0) {
      // enough CTAs for ~4 waves of 2 CTAs per SM, but never fewer tiles than warps per CTA
      const int64_t want = 148 * 2 * 4;
      nsplit = (int)std::max<int64_t>(1, div_up(want, B));
      const int ntile = cell_round_up(std::min(n1, n2), CELL_TILE) / CELL_TILE;
      nsplit = std::max(1, std::min(nsplit, ntile / (CELL_SWEEP_THREADS / 32)));
      // the symmetric sweep (half the pair work) needs one CTA per frame: worth a ragged last wave
      // as soon as there is about one CTA per SM
      float unused_scale;
      int unused_passes;
      if (B >= 148 && sym_plan<T, KIND, OVERLAP>(need_grad, 1, cell_round_up(n1, CELL_TILE), &unused_scale, &unused_passes))
        nsplit = 1;
    }
    // symmetric mode (one sweep for both groups): FP32 contacts with 1/(1+x^6), disjoint groups, one
    // CTA per frame, and a fixed-point scale of at least 2^16 that provably cannot overflow
    float sym_scale = 0.f;
    int sym_passes = 1;
    const bool sym = sym_plan<T, KIND, OVERLAP>(need_grad, nsplit, cell_round_up(n1, CELL_TILE), &sym_scale, &sym_passes);
    // memory layout of the sorted groups (CellLayout): a group that serves as the j side of a sweep
    // is dealt to that sweep's chunks; in symmetric mode the first group is only ever the i side
    const CellLayout lay1 = cell_layout(n1, sym ? 0 : cell_jchunk<T>(false));
    const CellLayout lay2 = cell_layout(n2, cell_jchunk<T>(sym));
    const int n1p = lay1.n_pad, n2p = lay2.n_pad;

    Carver carve(ws);
    T4* s1 = carve.take<T4>((size_t)B * n1p);
    T4* s2 = carve.take<T4>((size_t)B * n2p);
    T4* bb1 = carve.take<T4>((size_t)B * (n1p / CELL_CLUSTER) * 2);
    T4* bb2 = carve.take<T4>((size_t)B * (n2p / CELL_CLUSTER) * 2);
    T4* gr1 = need_grad ? carve.take<T4>((size_t)B * n1p) : nullptr;
    T4* gr2 = need_grad ? carve.take<T4>((size_t)B * n2p) : nullptr;
    double* partials = carve.take<double>((size_t)B * nsplit * NCH);
    double* excluded = carve.take<double>((size_t)B);
    using Acc = typename PairFn<T, KIND>::Acc;
    Acc* frame_coef = carve.take<Acc>((size_t)B * 2);
    const T4* prm_atom = static_cast<const T4*>(d_prm_atom[which].ptr);

    {
      const size_t smem = cell_sort_smem(std::max(n1, n2));
      auto kernel = cell_sort_kernel<T, CPBC>;
      CVP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      timer.begin("cell_sort", st);
      kernel<<<(unsigned)(2 * B), CELL_SORT_THREADS, smem, st>>>(
          pos, N, static_cast<const int32_t*>(d_g1.ptr), static_cast<const int32_t*>(d_tag1.ptr), n1,
          n1p, static_cast<const int32_t*>(d_g2.ptr), static_cast<const int32_t*>(d_tag2.ptr), n2,
          n2p, lay1.nchunk, lay2.nchunk, box, a.box_mode, s1, s2, bb1, bb2);
      timer.end(st);
      CVP_LAUNCH_CHECK();
    }
    const bool has_ex = !ex_pairs.empty() && KIND != CVP_PAIR_SOFTMIN;
    if (has_ex) {
      exclusion_energy_kernel<T, KIND, PBC><<<(unsigned)B, 128, 0, st>>>(
          pos, N, prm_atom, static_cast<const int32_t*>(d_ex_pairs.ptr),
          (int)(ex_pairs.size() / 2), box, a.box_mode, pc, excluded);
      CVP_LAUNCH_CHECK();
    }
    last_path = sym ? 2 : 1;
    // the symmetric sweep writes its gradients where they belong (no packed arrays, no scatter pass)
    const T gscale = KIND == CVP_PAIR_SOFTMIN ? T(1) : (T)d.scale;
    DirectGradient<T> direct;
    if (sym) {
      direct.grad = grad;
      direct.num_atoms = N;
      direct.scale = gscale;
      direct.accumulate = accumulate ? 1 : 0;
      if (!accumulate) CVP_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)B * N * 3 * sizeof(T), st));
    }
    auto sweep = [&](bool first, bool with_grad, const Acc* coef, double* energy) -> int {
      const T4* ip = first ? s1 : s2;
      const T4* jp = first ? s2 : s1;
      const T4* jb = first ? bb2 : bb1;
      const int nip = first ? n1p : n2p, njp = first ? n2p : n1p;
      const CellLayout& ilay = first ? lay1 : lay2;
      const CellLayout& jlay = first ? lay2 : lay1;
      T4* out = first ? gr1 : gr2;
      // default step function 1/(1+x^6): specialised pair function
      const bool fast6 = KIND == CVP_PAIR_CONTACTS && d.step_kind == CVP_STEP_RATIONAL &&
                         (int)d.step_p0 == 6;
      constexpr bool CAN_FAST6 = KIND == CVP_PAIR_CONTACTS;
      auto kernel = with_grad ? cell_sweep_kernel<T, KIND, CPBC, OVERLAP, true, false>
                              : cell_sweep_kernel<T, KIND, CPBC, OVERLAP, false, false>;
      if (fast6)
        kernel = with_grad ? cell_sweep_kernel<T, KIND, CPBC, OVERLAP, true, CAN_FAST6>
                           : cell_sweep_kernel<T, KIND, CPBC, OVERLAP, false, CAN_FAST6>;
      size_t sweep_smem = cell_sweep_smem<T>(false);
      T4* jout = nullptr;
      if constexpr (std::is_same<T, float>::value && KIND == CVP_PAIR_CONTACTS && !OVERLAP) {
        if (sym && first && with_grad) {
          kernel = cell_sweep_kernel<T, KIND, CPBC, OVERLAP, true, true, true>;
          sweep_smem = cell_sweep_smem<T>(true);
          jout = gr2;
        }
      }
      CVP_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sweep_smem));
      timer.begin("sweep", st);
      kernel<<<(unsigned)(B * nsplit), CELL_SWEEP_THREADS, sweep_smem, st>>>(
          ip, jp, jb, prm_atom, nip, njp, nsplit, box, a.box_mode, pc, coef, energy,
          with_grad ? out : nullptr, jout, sym_scale, sym_passes, jlay.cpc * CELL_CLUSTER, ilay.nchunk,
          jout != nullptr ? direct : DirectGradient<T>{});
      timer.end(st);
      CVP_LAUNCH_CHECK();
      return CVP_OK;
    };
    const double rc = d.cutoff;
    const double exp_neg_beta = exp(-d.beta);
    const int finish_blocks = (int)div_up(B * 32, 128);
    int status;
    if (KIND == CVP_PAIR_SOFTMIN) {
      if ((status = sweep(true, false, nullptr, partials)) != CVP_OK) return status;
      pair_finish_kernel<T, KIND><<<finish_blocks, 128, 0, st>>>(
          partials, nsplit, nullptr, 1.0, rc, exp_neg_beta, value, frame_coef, accumulate ? 1 : 0, B);
      CVP_LAUNCH_CHECK();
      if (need_grad && (status = sweep(true, true, frame_coef, nullptr)) != CVP_OK) return status;
    } else {
      if ((status = sweep(true, need_grad, nullptr, partials)) != CVP_OK) return status;
      pair_finish_kernel<T, KIND><<<finish_blocks, 128, 0, st>>>(
          partials, nsplit, has_ex ? excluded : nullptr, d.scale, rc, exp_neg_beta, value, nullptr,
          accumulate ? 1 : 0, B);
      CVP_LAUNCH_CHECK();
    }
    if (!need_grad) return CVP_OK;
    if (!sym && (status = sweep(false, true, frame_coef, nullptr)) != CVP_OK) return status;
    if (!sym) {
      if (!accumulate) CVP_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)B * N * 3 * sizeof(T), st));
      int64_t total = B * (int64_t)n1p;
      int blocks = (int)std::min<int64_t>(div_up(total, 256), 148 * 16);
      cell_scatter_kernel<T><<<blocks, 256, 0, st>>>(s1, gr1, n1p, N, gscale, grad, total);
      CVP_LAUNCH_CHECK();
      total = B * (int64_t)n2p;
      blocks = (int)std::min<int64_t>(div_up(total, 256), 148 * 16);
      cell_scatter_kernel<T><<<blocks, 256, 0, st>>>(s2, gr2, n2p, N, gscale, grad, total);
      CVP_LAUNCH_CHECK();
    }
    if (has_ex) {
      int64_t total = B * (int64_t)ex_atoms.size();
      exclusion_gradient_kernel<T, KIND, PBC><<<(unsigned)div_up(total, 128), 128, 0, st>>>(
          pos, N, prm_atom, static_cast<const int32_t*>(d_ex_atoms.ptr),
          static_cast<const int32_t*>(d_ex_offsets.ptr),
          static_cast<const int32_t*>(d_ex_partners.ptr), (int)ex_atoms.size(), box, a.box_mode,
          pc, gscale, grad, total);
      CVP_LAUNCH_CHECK();
    }
    return CVP_OK;
  }

  template <typename T, int KIND, int PBC, bool OVERLAP>
  int run_impl(const EvalArgs& a, void* ws) {
    if (cells_eligible(PBC)) return run_cells<T, KIND, PBC, OVERLAP>(a, ws);
    last_path = 0;
    using T4 = typename Vec4<T>::type;
    constexpr int NCH = PairFn<T, KIND>::NCH;
    const int which = std::is_same<T, float>::value ? 0 : 1;
    const int64_t B = a.num_frames;
    const int64_t N = a.num_atoms;
    const int n1 = (int)g1.size(), n2 = (int)g2.size();
    const bool need_grad = a.grad != nullptr;
    const bool accumulate = (a.flags & CVP_FLAG_ACCUMULATE) != 0;
    const PairConsts<T> pc = consts<T>();
    const T* pos = static_cast<const T*>(a.positions);
    const T* box = static_cast<const T*>(a.box);
    T* value = static_cast<T*>(a.value);
    T* grad = static_cast<T*>(a.grad);
    cudaStream_t st = a.stream;

    Carver carve(ws);
    T4* p1 = carve.take<T4>((size_t)B * n1);
    T4* p2 = carve.take<T4>((size_t)B * n2);
    T4* gr1 = need_grad ? carve.take<T4>((size_t)B * n1) : nullptr;
    T4* gr2 = need_grad ? carve.take<T4>((size_t)B * n2) : nullptr;
    const int nb1 = nblk(n1), nb2 = nblk(n2);
    double* partials = carve.take<double>((size_t)B * nb1 * NCH);
    double* excluded = carve.take<double>((size_t)B);
    using Acc = typename PairFn<T, KIND>::Acc;
    Acc* frame_coef = carve.take<Acc>((size_t)B * 2);

    const T4* prm1 = static_cast<const T4*>(d_prm1[which].ptr);
    const T4* prm2 = static_cast<const T4*>(d_prm2[which].ptr);
    const T4* prm_atom = static_cast<const T4*>(d_prm_atom[which].ptr);
    const int32_t* i1 = static_cast<const int32_t*>(d_g1.ptr);
    const int32_t* i2 = static_cast<const int32_t*>(d_g2.ptr);

    {
      int64_t total = B * (int64_t)(n1 + n2);
      int blocks = (int)std::min<int64_t>(div_up(total, 256), 148 * 16);
      pack_groups_kernel<T><<<blocks, 256, 0, st>>>(
          pos, N, i1, static_cast<const int32_t*>(d_tag1.ptr), n1, i2,
          static_cast<const int32_t*>(d_tag2.ptr), n2, p1, p2, total);
      CVP_LAUNCH_CHECK();
    }

    const bool has_ex = !ex_pairs.empty() && KIND != CVP_PAIR_SOFTMIN;
    if (has_ex) {
      exclusion_energy_kernel<T, KIND, PBC><<<(unsigned)B, 128, 0, st>>>(
          pos, N, prm_atom, static_cast<const int32_t*>(d_ex_pairs.ptr),
          (int)(ex_pairs.size() / 2), box, a.box_mode, pc, excluded);
      CVP_LAUNCH_CHECK();
    }

    const double rc = d.cutoff;
    const double exp_neg_beta = exp(-d.beta);
    const unsigned grid1 = (unsigned)(B * nb1), grid2 = (unsigned)(B * nb2);
    const int finish_blocks = (int)div_up(B * 32, 128);

    // one sweep: `first` selects which group sits in registers (true: group1 over group2)
    auto sweep = [&](bool first, bool with_grad, const Acc* coef, double* energy) -> int {
      const T4* ip = first ? p1 : p2;
      const T4* jp = first ? p2 : p1;
      const T4* iq = first ? prm1 : prm2;
      const T4* jq = first ? prm2 : prm1;
      const int ni = first ? n1 : n2, nj = first ? n2 : n1, nbi = first ? nb1 : nb2;
      T4* out = first ? gr1 : gr2;
      timer.begin("sweep", st);
      if (with_grad)
        pair_sweep_kernel<T, KIND, PBC, OVERLAP, true><<<first ? grid1 : grid2, PAIR_THREADS, 0, st>>>(
            ip, jp, iq, jq, ni, nj, nbi, box, a.box_mode, pc, coef, energy, out);
      else
        pair_sweep_kernel<T, KIND, PBC, OVERLAP, false><<<first ? grid1 : grid2, PAIR_THREADS, 0, st>>>(
            ip, jp, iq, jq, ni, nj, nbi, box, a.box_mode, pc, coef, energy, nullptr);
      timer.end(st);
      CVP_LAUNCH_CHECK();
      return CVP_OK;
    };
    int status;
    if (KIND == CVP_PAIR_SOFTMIN) {
      // value pass first: the gradient of a ratio of sums needs the sums
      if ((status = sweep(true, false, nullptr, partials)) != CVP_OK) return status;
      pair_finish_kernel<T, KIND><<<finish_blocks, 128, 0, st>>>(
          partials, nb1, nullptr, 1.0, rc, exp_neg_beta, value, frame_coef, accumulate ? 1 : 0, B);
      CVP_LAUNCH_CHECK();
      if (need_grad && (status = sweep(true, true, frame_coef, nullptr)) != CVP_OK) return status;
    } else {
      if ((status = sweep(true, need_grad, nullptr, partials)) != CVP_OK) return status;
      pair_finish_kernel<T, KIND><<<finish_blocks, 128, 0, st>>>(
          partials, nb1, has_ex ? excluded : nullptr, d.scale, rc, exp_neg_beta, value, nullptr,
          accumulate ? 1 : 0, B);
      CVP_LAUNCH_CHECK();
    }
    if (!need_grad) return CVP_OK;
    if ((status = sweep(false, true, frame_coef, nullptr)) != CVP_OK) return status;

    if (!accumulate)
      CVP_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)B * N * 3 * sizeof(T), st));
    const T gscale = KIND == CVP_PAIR_SOFTMIN ? T(1) : (T)d.scale;
    {
      int64_t total = B * (int64_t)n1;
      int blocks = (int)std::min<int64_t>(div_up(total, 256), 148 * 16);
      scatter_add_kernel<T><<<blocks, 256, 0, st>>>(gr1, i1, n1, N, gscale, grad, total);
      CVP_LAUNCH_CHECK();
      total = B * (int64_t)n2;
      blocks = (int)std::min<int64_t>(div_up(total, 256), 148 * 16);
      scatter_add_kernel<T><<<blocks, 256, 0, st>>>(gr2, i2, n2, N, gscale, grad, total);
      CVP_LAUNCH_CHECK();
    }
    if (has_ex) {
      int64_t total = B * (int64_t)ex_atoms.size();
      exclusion_gradient_kernel<T, KIND, PBC><<<(unsigned)div_up(total, 128), 128, 0, st>>>(
          pos, N, prm_atom, static_cast<const int32_t*>(d_ex_atoms.ptr),
          static_cast<const int32_t*>(d_ex_offsets.ptr),
          static_cast<const int32_t*>(d_ex_partners.ptr), (int)ex_atoms.size(), box, a.box_mode,
          pc, gscale, grad, total);
      CVP_LAUNCH_CHECK();
    }
    return CVP_OK;
  }

  template <typename T, int KIND>
  int run_kind(const EvalArgs& a, void* ws) {
    int pbc = 0;
    if (d.periodic) {
      CVP_REQUIRE(a.box != nullptr && a.box_mode != CVP_BOX_NONE,
                  "this collective variable uses periodic boundary conditions but no box was given");
      pbc = (a.flags & CVP_FLAG_TRICLINIC) ? 2 : 1;
    }
#define CVP_PAIR_DISPATCH(P)                                                  \
  return overlap ? run_impl<T, KIND, P, true>(a, ws) : run_impl<T, KIND, P, false>(a, ws)
    if (pbc == 0) CVP_PAIR_DISPATCH(0);
    if (pbc == 1) CVP_PAIR_DISPATCH(1);
    CVP_PAIR_DISPATCH(2);
#undef CVP_PAIR_DISPATCH
  }

  template <typename T>
  int run_type(const EvalArgs& a, void* ws) {
    switch (d.kind) {
      case CVP_PAIR_CONTACTS:
        return run_kind<T, CVP_PAIR_CONTACTS>(a, ws);
      case CVP_PAIR_ATTRACTION:
        return run_kind<T, CVP_PAIR_ATTRACTION>(a, ws);
      default:
        return run_kind<T, CVP_PAIR_SOFTMIN>(a, ws);
    }
  }

  int run(const EvalArgs& a, void* ws) override {
    return a.dtype == CVP_F32 ? run_type<float>(a, ws) : run_type<double>(a, ws);
  }
};

}  // namespace

int pair_sum_create(const cvp_pair_sum_desc* desc, cvp_cv** out) {
  CVP_REQUIRE(desc != nullptr && out != nullptr, "null argument");
  CVP_REQUIRE(desc->kind >= CVP_PAIR_CONTACTS && desc->kind <= CVP_PAIR_SOFTMIN, "unknown pair kind");
  CVP_REQUIRE(desc->num_atoms > 0 && desc->num_atoms <= TAG_INDEX_MASK, "bad num_atoms");
  CVP_REQUIRE(desc->n1 > 0 && desc->n2 > 0 && desc->group1 && desc->group2, "empty group");
  if (desc->kind == CVP_PAIR_CONTACTS) {
    CVP_REQUIRE(desc->r0 > 0, "threshold distance must be positive");
    CVP_REQUIRE(desc->step_kind >= CVP_STEP_RATIONAL && desc->step_kind <= CVP_STEP_PLUMED,
                "unknown step function");
    CVP_REQUIRE(desc->step_kind == CVP_STEP_HEAVISIDE || desc->step_p0 >= 1,
                "step-function exponent must be >= 1");
  }
  if (desc->kind == CVP_PAIR_ATTRACTION)
    CVP_REQUIRE(desc->charge && desc->sigma && desc->epsilon && desc->cutoff > 0,
                "attraction strength needs per-atom parameters and a cutoff");
  if (desc->kind == CVP_PAIR_SOFTMIN)
    CVP_REQUIRE(desc->cutoff > 0, "shortest distance needs a cutoff");

  auto cv = new PairSumCV();
  cv->d = *desc;
  cv->num_atoms = desc->num_atoms;
  // interaction groups are sets: sort, drop duplicates
  auto as_set = [&](const int32_t* p, int n, std::vector<int32_t>& v) -> bool {
    v.assign(p, p + n);
    for (int32_t a : v)
      if (a < 0 || a >= desc->num_atoms) return false;
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    return true;
  };
  if (!as_set(desc->group1, desc->n1, cv->g1) || !as_set(desc->group2, desc->n2, cv->g2)) {
    delete cv;
    set_error("atom index out of range in a group");
    return CVP_ERR_INVALID;
  }
  std::vector<char> in1(desc->num_atoms, 0), in2(desc->num_atoms, 0);
  for (int32_t a : cv->g1) in1[a] = 1;
  for (int32_t a : cv->g2) in2[a] = 1;
  size_t both = 0;
  for (int a = 0; a < desc->num_atoms; ++a) both += (in1[a] && in2[a]);
  cv->overlap = both > 0;
  cv->identical = both == cv->g1.size() && both == cv->g2.size();
  auto tags = [&](const std::vector<int32_t>& g, std::vector<int32_t>& tag) {
    tag.resize(g.size());
    for (size_t k = 0; k < g.size(); ++k)
      tag[k] = g[k] | ((in1[g[k]] && in2[g[k]]) ? TAG_BOTH : 0);
  };
  tags(cv->g1, cv->tag1);
  tags(cv->g2, cv->tag2);
  // exclusions that can occur in the interaction group, unique unordered pairs
  {
    std::vector<std::pair<int32_t, int32_t>> pairs;
    for (int e = 0; e < desc->num_exclusions; ++e) {
      int32_t a = desc->exclusions[2 * e], b = desc->exclusions[2 * e + 1];
      if (a < 0 || b < 0 || a >= desc->num_atoms || b >= desc->num_atoms) {
        delete cv;
        set_error("atom index out of range in an exclusion");
        return CVP_ERR_INVALID;
      }
      if (a == b) continue;
      if (!((in1[a] && in2[b]) || (in1[b] && in2[a]))) continue;
      pairs.emplace_back(std::min(a, b), std::max(a, b));
    }
    std::sort(pairs.begin(), pairs.end());
    pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
    std::vector<std::vector<int32_t>> partners(desc->num_atoms);
    for (auto& pr : pairs) {
      cv->ex_pairs.push_back(pr.first);
      cv->ex_pairs.push_back(pr.second);
      partners[pr.first].push_back(pr.second);
      partners[pr.second].push_back(pr.first);
    }
    cv->ex_offsets.push_back(0);
    for (int a = 0; a < desc->num_atoms; ++a) {
      if (partners[a].empty()) continue;
      cv->ex_atoms.push_back(a);
      for (int32_t p : partners[a]) cv->ex_partners.push_back(p);
      cv->ex_offsets.push_back((int32_t)cv->ex_partners.size());
    }
  }
  if (desc->kind == CVP_PAIR_ATTRACTION) {
    cv->charge.assign(desc->charge, desc->charge + desc->num_atoms);
    cv->sigma.assign(desc->sigma, desc->sigma + desc->num_atoms);
    cv->epsilon.assign(desc->epsilon, desc->epsilon + desc->num_atoms);
    if (desc->sign) cv->sign.assign(desc->sign, desc->sign + desc->num_atoms);
  }
  // the spec keeps its own copies
  cv->d.group1 = cv->d.group2 = cv->d.exclusions = nullptr;
  cv->d.charge = cv->d.sigma = cv->d.epsilon = cv->d.sign = nullptr;
  for (int a = 0; a < desc->num_atoms; ++a)
    if (in1[a] || in2[a]) cv->active_atoms.push_back(a);
  *out = cv;
  return CVP_OK;
}


// ================================================================================================
// Centroid pair sums: ResidueCoordination (cvpack/residue_coordination.py:129-144).
// Centroids R = sum_a w_a r_a are formed from raw coordinates, then the same sweep kernels run over
// the two sets of centroids (no cutoff, minimum image on R_J - R_I only), and the centroid
// gradients are handed down to the atoms: d/dr_a = w_a d/dR.
// ================================================================================================
namespace {

template <typename T>
__global__ void centroid_kernel(const T* __restrict__ pos, int64_t num_atoms,
                                const int32_t* __restrict__ offsets,
                                const int32_t* __restrict__ atoms,
                                const double* __restrict__ weights, int n1, int n2,
                                typename Vec4<T>::type* __restrict__ out1,
                                typename Vec4<T>::type* __restrict__ out2, int64_t total) {
  using T4 = typename Vec4<T>::type;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int n = n1 + n2;
  const int64_t b = t / n;
  const int k = (int)(t - b * n);
  const T* frame = pos + b * num_atoms * 3;
  double cx = 0, cy = 0, cz = 0;
  for (int e = offsets[k]; e < offsets[k + 1]; ++e) {
    const int64_t atom = atoms[e];
    const double w = weights[e];
    cx += w * frame[3 * atom];
    cy += w * frame[3 * atom + 1];
    cz += w * frame[3 * atom + 2];
  }
  T4 rec;
  rec.x = T(cx);
  rec.y = T(cy);
  rec.z = T(cz);
  rec.w = tag_encode(k, T(0));
  if (k < n1)
    out1[b * n1 + k] = rec;
  else
    out2[b * n2 + (k - n1)] = rec;
}

// one thread per (frame, active atom): sum over the centroid groups the atom belongs to
template <typename T>
__global__ void centroid_scatter_kernel(const typename Vec4<T>::type* __restrict__ g1,
                                        const typename Vec4<T>::type* __restrict__ g2, int n1,
                                        int n2, const int32_t* __restrict__ active_atom,
                                        const int32_t* __restrict__ atom_offsets,
                                        const int32_t* __restrict__ atom_group,
                                        const double* __restrict__ atom_weight, int num_active,
                                        int64_t num_atoms, T scale, T* __restrict__ grad,
                                        int accumulate, int64_t total) {
  using T4 = typename Vec4<T>::type;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t b = t / num_active;
  const int k = (int)(t - b * num_active);
  T gx = 0, gy = 0, gz = 0;
  for (int m = atom_offsets[k]; m < atom_offsets[k + 1]; ++m) {
    const int grp = atom_group[m];
    const T w = scale * T(atom_weight[m]);
    const T4 g = grp < n1 ? g1[b * n1 + grp] : g2[b * n2 + (grp - n1)];
    gx += w * g.x;
    gy += w * g.y;
    gz += w * g.z;
  }
  T* out = grad + (b * num_atoms + active_atom[k]) * 3;
  if (accumulate) {
    out[0] += gx;
    out[1] += gy;
    out[2] += gz;
  } else {
    out[0] = gx;
    out[1] = gy;
    out[2] = gz;
  }
}

struct CentroidPairSumCV : cvp_cv {
  cvp_centroid_pair_sum_desc d{};
  std::vector<int32_t> offsets, atoms;
  std::vector<double> weights;
  std::vector<int32_t> atom_offsets, atom_group;
  std::vector<double> atom_weight;
  DeviceBuffer d_offsets, d_atoms, d_weights, d_active, d_atom_offsets, d_atom_group,
      d_atom_weight;
  ~CentroidPairSumCV() override {
    for (DeviceBuffer* buf : {&d_offsets, &d_atoms, &d_weights, &d_active, &d_atom_offsets,
                              &d_atom_group, &d_atom_weight})
      buf->release();
  }
  int upload() override {
    int status;
    if ((status = cvp::upload(d_offsets, offsets)) != CVP_OK) return status;
    if ((status = cvp::upload(d_atoms, atoms)) != CVP_OK) return status;
    if ((status = cvp::upload(d_weights, weights)) != CVP_OK) return status;
    if ((status = cvp::upload(d_active, active_atoms)) != CVP_OK) return status;
    if ((status = cvp::upload(d_atom_offsets, atom_offsets)) != CVP_OK) return status;
    if ((status = cvp::upload(d_atom_group, atom_group)) != CVP_OK) return status;
    return cvp::upload(d_atom_weight, atom_weight);
  }
  int nblk(int n) const { return (int)div_up(n, PAIR_I_PER_BLOCK); }
  size_t workspace_per_frame(int dtype, bool need_grad) const override {
    size_t rec = dtype == CVP_F32 ? 16 : 32;
    size_t n = (size_t)d.n1 + d.n2;
    return n * rec * (need_grad ? 2 : 1) + 1024 + (size_t)nblk(d.n1) * sizeof(double) + 64;
  }
  size_t workspace_fixed(int) const override { return 8 * 256; }
  int64_t max_frames_per_run() const override {
    int64_t blocks = std::max(nblk(d.n1), nblk(d.n2));
    return std::max<int64_t>(1, (int64_t(1) << 31) / blocks - 1);
  }

  template <typename T, int PBC>
  int run_impl(const EvalArgs& a, void* ws) {
    using T4 = typename Vec4<T>::type;
    constexpr int KIND = CVP_PAIR_CONTACTS;
    const int64_t B = a.num_frames, N = a.num_atoms;
    const int n1 = d.n1, n2 = d.n2;
    const bool need_grad = a.grad != nullptr;
    const bool accumulate = (a.flags & CVP_FLAG_ACCUMULATE) != 0;
    const T* pos = static_cast<const T*>(a.positions);
    const T* box = static_cast<const T*>(a.box);
    T* value = static_cast<T*>(a.value);
    T* grad = static_cast<T*>(a.grad);
    cudaStream_t st = a.stream;
    PairConsts<T> pc{};
    pc.cutoff2 = (T)3.0e38;
    pc.has_switch = 0;
    pc.inv_r0 = (T)(1.0 / d.r0);
    pc.inv_r0sq = (T)(1.0 / (d.r0 * d.r0));
    pc.step.kind = d.step_kind;
    pc.step.n = (int)d.step_p0;
    pc.step.m = (int)d.step_p1;

    Carver carve(ws);
    T4* p1 = carve.take<T4>((size_t)B * n1);
    T4* p2 = carve.take<T4>((size_t)B * n2);
    T4* gr1 = need_grad ? carve.take<T4>((size_t)B * n1) : nullptr;
    T4* gr2 = need_grad ? carve.take<T4>((size_t)B * n2) : nullptr;
    const int nb1 = nblk(n1), nb2 = nblk(n2);
    double* partials = carve.take<double>((size_t)B * nb1);

    int64_t total = B * (int64_t)(n1 + n2);
    centroid_kernel<T><<<(unsigned)div_up(total, 128), 128, 0, st>>>(
        pos, N, static_cast<const int32_t*>(d_offsets.ptr),
        static_cast<const int32_t*>(d_atoms.ptr), static_cast<const double*>(d_weights.ptr), n1,
        n2, p1, p2, total);
    CVP_LAUNCH_CHECK();
    const unsigned grid1 = (unsigned)(B * nb1), grid2 = (unsigned)(B * nb2);
    // the default step function 1/(1+x^6) has a pair function that needs r^2 only
    const bool fast6 = d.step_kind == CVP_STEP_RATIONAL && (int)d.step_p0 == 6;
    auto value_kernel = fast6 ? pair_sweep_kernel<T, KIND, PBC, false, false, true>
                              : pair_sweep_kernel<T, KIND, PBC, false, false, false>;
    auto grad_kernel = fast6 ? pair_sweep_kernel<T, KIND, PBC, false, true, true>
                             : pair_sweep_kernel<T, KIND, PBC, false, true, false>;
    timer.begin("sweep", st);
    if (need_grad)
      grad_kernel<<<grid1, PAIR_THREADS, 0, st>>>(
          p1, p2, nullptr, nullptr, n1, n2, nb1, box, a.box_mode, pc, nullptr, partials, gr1);
    else
      value_kernel<<<grid1, PAIR_THREADS, 0, st>>>(
          p1, p2, nullptr, nullptr, n1, n2, nb1, box, a.box_mode, pc, nullptr, partials, nullptr);
    timer.end(st);
    CVP_LAUNCH_CHECK();
    pair_finish_kernel<T, KIND><<<(unsigned)div_up(B * 32, 128), 128, 0, st>>>(
        partials, nb1, nullptr, d.scale, 0.0, 0.0, value, nullptr, accumulate ? 1 : 0, B);
    CVP_LAUNCH_CHECK();
    if (!need_grad) return CVP_OK;
    timer.begin("sweep", st);
    grad_kernel<<<grid2, PAIR_THREADS, 0, st>>>(
        p2, p1, nullptr, nullptr, n2, n1, nb2, box, a.box_mode, pc, nullptr, nullptr, gr2);
    timer.end(st);
    CVP_LAUNCH_CHECK();
    if (!accumulate) CVP_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)B * N * 3 * sizeof(T), st));
    const int num_active = (int)active_atoms.size();
    total = B * (int64_t)num_active;
    centroid_scatter_kernel<T><<<(unsigned)div_up(total, 128), 128, 0, st>>>(
        gr1, gr2, n1, n2, static_cast<const int32_t*>(d_active.ptr),
        static_cast<const int32_t*>(d_atom_offsets.ptr),
        static_cast<const int32_t*>(d_atom_group.ptr),
        static_cast<const double*>(d_atom_weight.ptr), num_active, N, (T)d.scale, grad,
        accumulate ? 1 : 0, total);
    CVP_LAUNCH_CHECK();
    return CVP_OK;
  }
  template <typename T>
  int run_type(const EvalArgs& a, void* ws) {
    if (!d.periodic) return run_impl<T, 0>(a, ws);
    CVP_REQUIRE(a.box != nullptr && a.box_mode != CVP_BOX_NONE,
                "this collective variable uses periodic boundary conditions but no box was given");
    return (a.flags & CVP_FLAG_TRICLINIC) ? run_impl<T, 2>(a, ws) : run_impl<T, 1>(a, ws);
  }
  int run(const EvalArgs& a, void* ws) override {
    return a.dtype == CVP_F32 ? run_type<float>(a, ws) : run_type<double>(a, ws);
  }
};

}  // namespace

int centroid_pair_sum_create(const cvp_centroid_pair_sum_desc* desc, cvp_cv** out) {
  CVP_REQUIRE(desc != nullptr && out != nullptr, "null argument");
  CVP_REQUIRE(desc->num_atoms > 0 && desc->n1 > 0 && desc->n2 > 0 && desc->group_offsets &&
                  desc->group_atoms && desc->group_weights,
              "empty residue group");
  CVP_REQUIRE(desc->r0 > 0, "threshold distance must be positive");
  CVP_REQUIRE(desc->step_kind >= CVP_STEP_RATIONAL && desc->step_kind <= CVP_STEP_PLUMED,
              "unknown step function");
  const int n = desc->n1 + desc->n2;
  CVP_REQUIRE(desc->group_offsets[0] == 0, "bad group offsets");
  auto cv = new CentroidPairSumCV();
  cv->d = *desc;
  cv->num_atoms = desc->num_atoms;
  cv->offsets.assign(desc->group_offsets, desc->group_offsets + n + 1);
  const int total = cv->offsets[n];
  cv->atoms.assign(desc->group_atoms, desc->group_atoms + total);
  cv->weights.assign(desc->group_weights, desc->group_weights + total);
  cv->d.group_offsets = cv->d.group_atoms = nullptr;
  cv->d.group_weights = nullptr;
  std::vector<std::vector<std::pair<int32_t, double>>> by_atom(desc->num_atoms);
  for (int g = 0; g < n; ++g) {
    if (cv->offsets[g + 1] <= cv->offsets[g]) {
      delete cv;
      set_error("a residue has no atoms");
      return CVP_ERR_INVALID;
    }
    for (int e = cv->offsets[g]; e < cv->offsets[g + 1]; ++e) {
      int32_t atom = cv->atoms[e];
      if (atom < 0 || atom >= desc->num_atoms) {
        delete cv;
        set_error("atom index out of range in a residue");
        return CVP_ERR_INVALID;
      }
      by_atom[atom].emplace_back(g, cv->weights[e]);
    }
  }
  cv->atom_offsets.push_back(0);
  for (int a = 0; a < desc->num_atoms; ++a) {
    if (by_atom[a].empty()) continue;
    cv->active_atoms.push_back(a);
    for (auto& m : by_atom[a]) {
      cv->atom_group.push_back(m.first);
      cv->atom_weight.push_back(m.second);
    }
    cv->atom_offsets.push_back((int32_t)cv->atom_group.size());
  }
  *out = cv;
  return CVP_OK;
}

}  // namespace cvp
