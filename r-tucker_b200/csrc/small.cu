// (c) The N-independent part of one Riemannian optimiser step, in fp64 on the device.
//
// Restates, at rank r (no rank-2r construct, no autodiff), what the reference obtains from
// tucker_riemopt 1.0.1: the (S_(i)S_(i)^T)^-1 factors and the gauge projection of `grad`
// (call sites src/model/asymmetric/optim.py:89, symmetric/optim.py:83), `TangentVector.norm`
// (asymmetric/optim.py:90), `project` (asymmetric/optim.py:86) and the core part of
// `construct().round(rank)` (asymmetric/optim.py:108) -- the latter as the structured HOSVD of
// SURVEY.md App. A.5 (Cholesky of the W_i Grams, block-structured rank-2r core, eigenvectors of the
// unfolding Grams by block Jacobi) instead of QR(N x 2r) + SVD(2r_i x prod 2r_j).
//
// Everything here costs O(poly(r)) and is replicated on every GPU of an entity-sharded run.
#include "small_kernels.cuh"

#include <cstdlib>
// Stop criterion of the HOSVD eigenproblems (off-diagonal mass SEEN in a sweep, relative to ||A||_F^2).
// RT_HOSVD_STOP overrides it for experiments.
static double hosvd_stop() {
  static const double v = [] { const char* e = std::getenv("RT_HOSVD_STOP"); return e ? std::atof(e) : 1e-10; }();
  return v;
}

namespace rt {
int eig_batch(int count, const double* const* A, const int* n, double* const* w, double* const* V,
              void* const* ws, cudaStream_t s, double stop);
size_t eig_ws_bytes(int n);
int subspace_batch(int count, const double* const* N, const int* n, const int* r, double* const* Y, const int* ldy,
                   void* const* ws, void* shared_ws, int* const* info, cudaStream_t s);
size_t subspace_ws_bytes(int n, int r);
size_t subspace_shared_ws_bytes();
}  // namespace rt

// HOSVD route: 1 (default) = dominant subspaces by purification + Newton-Schulz on the fp64 tensor cores
// (subspace.cu); 0 = full eigen-decomposition by block Jacobi (eig.cu, round 1).  RT_HOSVD=jacobi selects 0.
static int hosvd_route() {
  static const int v = [] { const char* e = std::getenv("RT_HOSVD"); return (e && e[0] == 'j') ? 0 : 1; }();
  return v;
}

extern "C" size_t rt_gram_ws_bytes(int n, int ra, int rb);

namespace {
using namespace rt::small;
using rt::align_up;
using rt::cdiv;

constexpr int kNumT = 8;          // core-sized fp64 temporaries
constexpr int kMaxSplitCtas = 592;
constexpr int kDotBlocks = 296;

struct Layout {
  int r[3];
  int B;
  int64_t c;  // r0*r1*r2
  size_t C64, Gm[3], Ainv[3], coresq, Tn[kNumT], KC[3], Nn[3], Vf[3], Wv[3], GL[3], GLinv[3], Gs[3],
      tmpM[3], tmpK[3], eig[3], sub[3], sub_shared, gemm_partial, gram_ws, dot_partial, scal, total;
};

Layout make_layout(int r0, int r1, int r2, int B) {
  Layout L;
  L.r[0] = r0; L.r[1] = r1; L.r[2] = r2; L.B = B;
  L.c = (int64_t)r0 * r1 * r2;
  size_t o = 0;
  auto take = [&](size_t doubles) { size_t at = o; o += align_up(doubles * sizeof(double), 256); return at; };
  L.C64 = take(L.c);
  for (int i = 0; i < 3; ++i) { L.Gm[i] = take((size_t)L.r[i] * L.r[i]); L.Ainv[i] = take((size_t)L.r[i] * L.r[i]); }
  L.coresq = take(4);
  for (int t = 0; t < kNumT; ++t) L.Tn[t] = take(L.c);
  for (int i = 0; i < 3; ++i) {
    const size_t r = L.r[i];
    L.KC[i] = take(2 * r * r);
    L.Nn[i] = take(4 * r * r);
    L.Vf[i] = take(4 * r * r);
    L.Wv[i] = take(2 * r);
    L.GL[i] = take(r * r);
    L.GLinv[i] = take(r * r);
    L.Gs[i] = take(r * r);
    L.tmpM[i] = take(2 * r * r);
    L.tmpK[i] = take(2 * r * r);
    size_t at = o; o += align_up(rt::eig_ws_bytes(2 * (int)r), 256); L.eig[i] = at;
    at = o; o += align_up(rt::subspace_ws_bytes(2 * (int)r, (int)r), 256); L.sub[i] = at;
  }
  { size_t at = o; o += align_up(rt::subspace_shared_ws_bytes(), 256); L.sub_shared = at; }
  L.gemm_partial = take((size_t)kMaxSplitCtas * GT * GT);
  int rmax = r0 > r1 ? r0 : r1; rmax = rmax > r2 ? rmax : r2;
  { size_t at = o; o += align_up(rt_gram_ws_bytes(B > 0 ? B : 1, rmax, rmax), 256); L.gram_ws = at; }
  L.dot_partial = take(kDotBlocks);
  L.scal = take(16);
  L.total = o;
  return L;
}

struct Ctx {
  Layout L;
  char* base;
  cudaStream_t s;
  int err = 0;
  double* p(size_t off) const { return (double*)(base + off); }
  double* T(int t) const { return p(L.Tn[t]); }
  int r(int i) const { return L.r[i]; }

  template <typename T>
  void gemm(GemmT<T> g) {
    if (err) return;
    if (g.m <= 0 || g.n <= 0) return;
    g.ksplit = 1; g.k_per_split = g.K1 * g.K2; g.partial = reinterpret_cast<T*>(p(L.gemm_partial));
    const int tiles = cdiv(g.m, GT) * cdiv(g.n, GT);
    const int K = g.K1 * g.K2;
    if (g.batch == 1 && tiles < 148 && K >= 8 * GK) {
      int ks = cdiv(296, tiles);
      const int max_ks = cdiv(K, 4 * GK);
      if (ks > max_ks) ks = max_ks;
      if ((int64_t)ks * tiles > kMaxSplitCtas) ks = kMaxSplitCtas / tiles;
      if (ks > 1) {
        g.k_per_split = cdiv(cdiv(K, ks), GK) * GK;
        g.ksplit = cdiv(K, g.k_per_split);
      }
    }
    dim3 grid(cdiv(g.m, GT), cdiv(g.n, GT), g.ksplit > 1 ? g.ksplit : g.batch);
    gemm64_kernel<T><<<grid, 256, 0, s>>>(g);
    ++rt::g_launches;
    if (g.ksplit > 1) {
      if (g.ksplit >= 64 && g.m * g.n <= 4096) gemm64_reduce_kernel<T, 32><<<cdiv(g.m * g.n * 32, 256), 256, 0, s>>>(g);
      else gemm64_reduce_kernel<T, 1><<<cdiv(g.m * g.n, 256), 256, 0, s>>>(g);
      ++rt::g_launches;
    }
    if (cudaGetLastError() != cudaSuccess) err = 1;
  }

  // Y = alpha * (X x_mode M) + beta * Y.  M is (mo x mi) with strides (sm_o, sm_i);
  // X has dims d[] with d[mode] == mi; Y has the same dims except d[mode] -> mo.
  template <typename T>
  void mode_prod(int mode, const T* M, int64_t sm_o, int64_t sm_i, int mo, int mi,
                 const T* X, const int d[3], T* Y, double alpha, double beta) {
    GemmT<T> g{};
    g.alpha = alpha; g.beta = beta; g.batch = 1;
    if (mode == 0) {
      const int64_t rest = (int64_t)d[1] * d[2];
      g.A = M; g.B = X; g.C = Y; g.m = mo; g.n = (int)rest; g.K1 = 1; g.K2 = mi;
      g.a_m = sm_o; g.a_k1 = 0; g.a_k2 = sm_i; g.b_k1 = 0; g.b_k2 = rest; g.b_n = 1;
      g.c_m = rest; g.c_n = 1;
    } else if (mode == 1) {
      g.A = M; g.B = X; g.C = Y; g.m = mo; g.n = d[2]; g.K1 = 1; g.K2 = mi;
      g.a_m = sm_o; g.a_k1 = 0; g.a_k2 = sm_i; g.b_k1 = 0; g.b_k2 = d[2]; g.b_n = 1;
      g.c_m = d[2]; g.c_n = 1;
      g.batch = d[0]; g.a_b = 0; g.b_b = (int64_t)mi * d[2]; g.c_b = (int64_t)mo * d[2];
    } else {
      g.A = X; g.B = M; g.C = Y; g.m = d[0] * d[1]; g.n = mo; g.K1 = 1; g.K2 = mi;
      g.a_m = mi; g.a_k1 = 0; g.a_k2 = 1; g.b_k1 = 0; g.b_k2 = sm_i; g.b_n = sm_o;
      g.c_m = mo; g.c_n = 1;
    }
    gemm(g);
  }

  // out[mx, my] (row stride ldo) = alpha * X_(mode) Y_(mode)^T + beta * out.
  // X dims dx[], Y dims equal to dx except dy_mode in the contracted-free mode.
  template <typename T>
  void unfold_gram(int mode, const T* X, const int dx[3], const T* Y, int my, T* out,
                   int64_t ldo, double alpha, double beta) {
    GemmT<T> g{};
    g.alpha = alpha; g.beta = beta; g.batch = 1;
    g.A = X; g.B = Y; g.C = out; g.m = dx[mode]; g.n = my; g.c_m = ldo; g.c_n = 1;
    if (mode == 0) {
      const int64_t rest = (int64_t)dx[1] * dx[2];
      g.K1 = 1; g.K2 = (int)rest;
      g.a_m = rest; g.a_k1 = 0; g.a_k2 = 1; g.b_k1 = 0; g.b_k2 = 1; g.b_n = rest;
    } else if (mode == 1) {
      g.K1 = dx[0]; g.K2 = dx[2];
      g.a_m = dx[2]; g.a_k1 = (int64_t)dx[1] * dx[2]; g.a_k2 = 1;
      g.b_n = dx[2]; g.b_k1 = (int64_t)my * dx[2]; g.b_k2 = 1;
    } else {
      g.K1 = 1; g.K2 = dx[0] * dx[1];
      g.a_m = 1; g.a_k1 = 0; g.a_k2 = dx[2]; g.b_k1 = 0; g.b_k2 = my; g.b_n = 1;
    }
    gemm(g);
  }

  // plain C[m,n] = alpha * A[m,k] B[k,n] + beta*C with row-major leading dims
  void matmul(const double* A, int64_t lda, bool ta, const double* B, int64_t ldb, bool tb, double* C,
              int64_t ldc, int m, int n, int k, double alpha, double beta) {
    Gemm g{};
    g.alpha = alpha; g.beta = beta; g.batch = 1;
    g.A = A; g.B = B; g.C = C; g.m = m; g.n = n; g.K1 = 1; g.K2 = k;
    g.a_m = ta ? 1 : lda; g.a_k2 = ta ? lda : 1;
    g.b_k2 = tb ? 1 : ldb; g.b_n = tb ? ldb : 1;
    g.c_m = ldc; g.c_n = 1;
    gemm(g);
  }

  float* Tf(int t) const { return reinterpret_cast<float*>(T(t)); }   // the same temporaries viewed as fp32
  void to32plain(const double* x, float* y, int64_t n) {
    if (err) return;
    int blocks = (int)((n + 255) / 256); if (blocks > 1184) blocks = 1184;
    f64_to_f32_kernel<<<blocks, 256, 0, s>>>(x, y, n); ++rt::g_launches;
  }
  void scale32(const float* x, float* y, int64_t n, double a_host, const double* a_dev) {
    if (err) return;
    int blocks = (int)((n + 255) / 256); if (blocks > 1184) blocks = 1184;
    scale_f32_kernel<<<blocks, 256, 0, s>>>(x, y, n, a_host, a_dev); ++rt::g_launches;
  }
  void to64(const float* x, double* y, int64_t n) {
    if (err) return;
    int blocks = (int)((n + 255) / 256); if (blocks > 1184) blocks = 1184;
    f32_to_f64_kernel<<<blocks, 256, 0, s>>>(x, y, n); ++rt::g_launches;
  }
  void to32(const double* x, float* y, int64_t n, double a_host, const double* a_dev) {
    if (err) return;
    int blocks = (int)((n + 255) / 256); if (blocks > 1184) blocks = 1184;
    f64_to_f32_scaled_kernel<<<blocks, 256, 0, s>>>(x, y, n, a_host, a_dev); ++rt::g_launches;
  }
  void axpby(const double* x, const double* y, double* z, int64_t n, double a, const double* a_dev,
             double b, const double* b_dev) {
    if (err) return;
    int blocks = (int)((n + 255) / 256); if (blocks > 1184) blocks = 1184;
    axpby64_kernel<<<blocks, 256, 0, s>>>(x, y, z, n, a, a_dev, b, b_dev); ++rt::g_launches;
  }
  template <typename TA, typename TB>
  void dot(const TA* x, const TB* y, int64_t n, double scale, double* out, int accumulate) {
    if (err) return;
    int blocks = (int)((n + 255) / 256); if (blocks > kDotBlocks) blocks = kDotBlocks; if (blocks < 1) blocks = 1;
    dot_partial_kernel<TA, TB><<<blocks, 256, 0, s>>>(x, y, n, p(L.dot_partial));
    dot_final_kernel<<<1, 32, 0, s>>>(p(L.dot_partial), blocks, scale, out, accumulate); rt::g_launches += 2;
  }
  int spd(const SpdBatch& b, int count, int nmax) {
    const size_t smem = ((size_t)nmax * (nmax + 1) / 2 + (size_t)SPD_NB * nmax) * sizeof(double);
    if (smem > 227 * 1024 || nmax > 256) {
      rt::set_error("small stage: rank %d exceeds the in-shared-memory Cholesky limit (232)", nmax); return 2; }
    static size_t configured = 0;
    if (smem > configured) {
      if (cudaFuncSetAttribute(spd_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        rt::set_error("small stage: cannot raise shared memory to %zu", smem); return 1; }
      configured = smem;
    }
    spd_blocked_kernel<<<count, SPD_THREADS, smem, s>>>(b); ++rt::g_launches;
    return 0;
  }
};

int finish(Ctx& c, const char* what) {
  if (c.err) { rt::set_error("%s: kernel launch failed", what); return 1; }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { rt::set_error("%s: %s", what, cudaGetErrorString(e)); return 1; }
  return 0;
}

// Contraction of a grouped rank-2r core with W_i = [W_ia | W_ib] (each r_i x r_i, given by strides):
//   out = X x(W0a,W1a,W2a) + B0 x(W0b,W1a,W2a) + B1 x(W0a,W1b,W2a) + B2 x(W0a,W1a,W2b)
// evaluated as   D  = X x0 W0a + B0 x0 W0b ;  E1 = B1 x0 W0a ;  E2 = B2 x0 W0a
//                U1 = D x1 W1a + E1 x1 W1b ;  U2 = E2 x1 W1a ;  out = U1 x2 W2a + U2 x2 W2b.
// Buffers are caller-provided; allowed aliases: E2 == E1 iff B2 == B1; U2 may reuse E1 when
// E2 != E1; out may reuse D.
template <typename T>
void grouped_contract(Ctx& c, const T* X, const T* const B[3], const T* const Wa[3],
                      const T* const Wb[3], const int64_t st_o[3], const int64_t st_i[3],
                      T* D, T* E1, T* E2, T* U1, T* U2, T* out) {
  const int d[3] = {c.r(0), c.r(1), c.r(2)};
  c.mode_prod(0, Wa[0], st_o[0], st_i[0], d[0], d[0], X, d, D, 1.0, 0.0);
  c.mode_prod(0, Wb[0], st_o[0], st_i[0], d[0], d[0], B[0], d, D, 1.0, 1.0);
  c.mode_prod(0, Wa[0], st_o[0], st_i[0], d[0], d[0], B[1], d, E1, 1.0, 0.0);
  if (E2 != E1) c.mode_prod(0, Wa[0], st_o[0], st_i[0], d[0], d[0], B[2], d, E2, 1.0, 0.0);
  c.mode_prod(1, Wa[1], st_o[1], st_i[1], d[1], d[1], D, d, U1, 1.0, 0.0);
  c.mode_prod(1, Wb[1], st_o[1], st_i[1], d[1], d[1], E1, d, U1, 1.0, 1.0);
  c.mode_prod(1, Wa[1], st_o[1], st_i[1], d[1], d[1], E2, d, U2, 1.0, 0.0);
  c.mode_prod(2, Wa[2], st_o[2], st_i[2], d[2], d[2], U1, d, out, 1.0, 0.0);
  c.mode_prod(2, Wb[2], st_o[2], st_i[2], d[2], d[2], U2, d, out, 1.0, 1.0);
}

}  // namespace

extern "C" size_t rt_small_ws_bytes(int r0, int r1, int r2, int B) {
  if (r0 <= 0 || r1 <= 0 || r2 <= 0) return 0;
  return make_layout(r0, r1, r2, B).total;
}

extern "C" int rt_small_prepare(const float* core, int r0, int r1, int r2, int sym, void* small_ws,
                                void* stream) {
  RT_REQUIRE(small_ws != nullptr && r0 > 0 && r1 > 0 && r2 > 0, "rt_small_prepare: bad arguments");
  RT_REQUIRE(!sym || r1 == r2, "rt_small_prepare: SF-Tucker needs r1 == r2");
  Ctx c{make_layout(r0, r1, r2, 0), (char*)small_ws, (cudaStream_t)stream};
  // layout must not depend on B for the persistent slots: they precede every B-dependent slot
  double* C64 = c.p(c.L.C64);
  c.to64(core, C64, c.L.c);
  const int d[3] = {r0, r1, r2};
  for (int i = 0; i < 3; ++i) c.unfold_gram(i, C64, d, C64, d[i], c.p(c.L.Gm[i]), d[i], 1.0, 0.0);
  if (sym) {
    c.axpby(c.p(c.L.Gm[1]), c.p(c.L.Gm[2]), c.p(c.L.Gm[1]), (int64_t)r1 * r1, 1.0, nullptr, 1.0, nullptr);
    c.axpby(c.p(c.L.Gm[1]), nullptr, c.p(c.L.Gm[2]), (int64_t)r1 * r1, 1.0, nullptr, 0.0, nullptr);
  }
  SpdBatch b{};
  int nmax = 0;
  for (int i = 0; i < 3; ++i) {
    b.p[i].G = c.p(c.L.Gm[i]); b.p[i].L = nullptr; b.p[i].Linv = nullptr; b.p[i].Ginv = c.p(c.L.Ainv[i]);
    b.p[i].n = d[i];
    nmax = d[i] > nmax ? d[i] : nmax;
  }
  int rc = c.spd(b, 3, nmax);
  if (rc) return rc;
  c.dot<double, double>(C64, C64, c.L.c, 1.0, c.p(c.L.coresq), 0);
  return finish(c, "rt_small_prepare");
}

extern "C" size_t rt_small_ainv_offset(int mode, int r0, int r1, int r2) {
  if (mode < 0 || mode > 2) return (size_t)-1;
  return make_layout(r0, r1, r2, 0).Ainv[mode];
}

extern "C" int rt_rows_times_ainv(const float* A, int m, int mode, int r0, int r1, int r2, float* C,
                                  void* small_ws, void* stream) {
  RT_REQUIRE(mode >= 0 && mode < 3 && m >= 0, "rt_rows_times_ainv: bad arguments");
  if (m == 0) return 0;
  Layout L = make_layout(r0, r1, r2, 0);
  const int r = L.r[mode];
  const double* K = (const double*)((char*)small_ws + L.Ainv[mode]);
  const size_t smem = (size_t)4 * r * sizeof(float);
  rows_times_mat_kernel<<<cdiv(m, 4), 256, smem, (cudaStream_t)stream>>>(A, m, r, K, C);
  RT_LAUNCH_CHECK();
  return 0;
}

extern "C" int rt_gram(const float* A, int64_t lda, const float* B, int64_t ldb, int n, int ra, int rb,
                       double* out, int precise, void* ws, void* stream);

extern "C" int rt_small_grad(const float* core, const float* d_core, const float* qp, const float* H,
                             const float* r_rows, const float* s_rows, const float* dr_rows,
                             const float* ds_rows, const double* bce_sum, double inv_count,
                             const double* hyper, int B, int r0, int r1, int r2, int sym, float* dS_g,
                             double* loss_total, float* drA, float* dsA, double* P_R, double* P_S,
                             double* P_O, void* small_ws, void* stream) {
  RT_REQUIRE(small_ws != nullptr && B > 0, "rt_small_grad: bad arguments");
  Ctx c{make_layout(r0, r1, r2, B), (char*)small_ws, (cudaStream_t)stream};
  cudaStream_t s = c.s;
  {
    int blocks = (int)((c.L.c + 255) / 256); if (blocks > 1184) blocks = 1184;
    grad_core_kernel<<<blocks, 256, 0, s>>>(d_core, core, hyper, dS_g, c.L.c); ++rt::g_launches;
  }
  // loss_total = bce_sum * inv_count + reg * ||core||^2
  c.axpby(bce_sum, c.p(c.L.coresq), loss_total, 1, inv_count, nullptr, 1.0, hyper + 1);
  int rc;
  if ((rc = rt_rows_times_ainv(dr_rows, B, 0, r0, r1, r2, drA, small_ws, stream))) return rc;
  if ((rc = rt_rows_times_ainv(ds_rows, B, 1, r0, r1, r2, dsA, small_ws, stream))) return rc;
  void* gws = c.base + c.L.gram_ws;
  // P_i = -(U_i^T g_i A_i): the Gram of the gathered rows with the A-scaled gradient rows
  if ((rc = rt_gram(r_rows, r0, drA, r0, B, r0, r0, P_R, 0, gws, stream))) return rc;
  c.axpby(P_R, nullptr, P_R, (int64_t)r0 * r0, -1.0, nullptr, 0.0, nullptr);
  if ((rc = rt_gram(s_rows, r1, dsA, r1, B, r1, r1, P_S, 0, gws, stream))) return rc;
  double* tmp = c.p(c.L.tmpK[2]);
  if ((rc = rt_gram(H, r2, qp, r2, B, r2, r2, sym ? tmp : P_O, 0, gws, stream))) return rc;
  if (sym) {
    c.axpby(P_S, tmp, P_S, (int64_t)r1 * r1, -1.0, nullptr, -1.0, nullptr);
    if (P_O && P_O != P_S) c.axpby(P_S, nullptr, P_O, (int64_t)r1 * r1, 1.0, nullptr, 0.0, nullptr);
  } else {
    c.axpby(P_S, nullptr, P_S, (int64_t)r1 * r1, -1.0, nullptr, 0.0, nullptr);
    c.axpby(P_O, nullptr, P_O, (int64_t)r2 * r2, -1.0, nullptr, 0.0, nullptr);
  }
  return finish(c, "rt_small_grad");
}

namespace {
__global__ void norm_finish_kernel(const double* sq, const double* hyper, double* norm_out,
                                   double* alpha_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double nrm = sqrt(fmax(sq[0], 0.0));
    norm_out[0] = nrm;
    const double ng = hyper[3];
    alpha_out[0] = (ng != 0.0) ? ng / nrm : 1.0;
  }
}
}  // namespace

extern "C" int rt_small_norm(const float* dS_g, const double* gram_R, const double* gram_S,
                             const double* gram_O, const double* hyper, int r0, int r1, int r2, int sym,
                             double* norm_out, double* alpha_out, void* small_ws, void* stream) {
  RT_REQUIRE(small_ws != nullptr, "rt_small_norm: workspace is NULL");
  Ctx c{make_layout(r0, r1, r2, 0), (char*)small_ws, (cudaStream_t)stream};
  double* sq = c.p(c.L.scal);
  c.dot<float, float>(dS_g, dS_g, c.L.c, 1.0, sq, 0);
  c.dot<double, double>(gram_R, c.p(c.L.Gm[0]), (int64_t)r0 * r0, 1.0, sq, 1);
  c.dot<double, double>(gram_S, c.p(c.L.Gm[1]), (int64_t)r1 * r1, 1.0, sq, 1);
  if (!sym) c.dot<double, double>(gram_O, c.p(c.L.Gm[2]), (int64_t)r2 * r2, 1.0, sq, 1);
  norm_finish_kernel<<<1, 32, 0, c.s>>>(sq, hyper, norm_out, alpha_out); ++rt::g_launches;
  return finish(c, "rt_small_norm");
}

extern "C" int rt_small_project(const float* core, const float* core_old, const float* dS_old,
                                const double* M_R, const double* M_S, const double* M_O,
                                const double* hyper, int r0, int r1, int r2, int sym, float* pS_beta,
                                double* K_R, double* K_S, double* K_O, double* L_R, double* L_S,
                                double* L_O, void* small_ws, void* stream) {
  RT_REQUIRE(small_ws != nullptr, "rt_small_project: workspace is NULL");
  (void)core;  // the current core is already resident as C64 (rt_small_prepare)
  Ctx c{make_layout(r0, r1, r2, 0), (char*)small_ws, (cudaStream_t)stream};
  const int d[3] = {r0, r1, r2};
  const double* M[3] = {M_R, M_S, sym ? M_S : M_O};
  double* Kout[3] = {K_R, K_S, K_O};
  double* Lout[3] = {L_R, L_S, L_O};
  // All O(r^4) contractions of the projection run in fp32 (inputs are the fp32 parameters themselves and
  // the result only needs fp32 accuracy); the r x r solves with (S_(i)S_(i)^T)^-1 stay in fp64.
  const float* Cf = core;
  const float* Co = core_old;
  const float* Xo = dS_old;
  float* Mf[3];
  for (int i = 0; i < 3; ++i) Mf[i] = reinterpret_cast<float*>(c.p(c.L.tmpM[i]));
  for (int i = 0; i < (sym ? 2 : 3); ++i) c.to32plain(M[i], Mf[i], 2 * (int64_t)d[i] * d[i]);
  if (sym) Mf[2] = Mf[1];
  const float* Wa[3]; const float* Wb[3]; int64_t st_o[3], st_i[3];
  for (int i = 0; i < 3; ++i) { Wa[i] = Mf[i]; Wb[i] = Mf[i] + d[i]; st_o[i] = 2 * d[i]; st_i[i] = 1; }
  const float* Cb[3] = {Co, Co, Co};
  float* D = c.Tf(2); float* Ta = c.Tf(3); float* U1 = c.Tf(4); float* U2 = c.Tf(5);
  float* pS = c.Tf(6);
  grouped_contract<float>(c, Xo, Cb, Wa, Wb, st_o, st_i, D, Ta, Ta, U1, U2, pS);
  c.scale32(pS, pS_beta, c.L.c, 1.0, hyper + 2);
  // KC_i (2 r_i x r_i) in fp32 scratch (the eigenvector slots are idle here), converted to fp64 afterwards
  float* KCf[3] = {reinterpret_cast<float*>(c.p(c.L.Vf[0])), reinterpret_cast<float*>(c.p(c.L.Vf[1])),
                   reinterpret_cast<float*>(c.p(c.L.Vf[2]))};
  double* KC[3] = {c.p(c.L.KC[0]), c.p(c.L.KC[1]), c.p(c.L.KC[2])};
  // KC_2 = [U1_(2) ; U2_(2)] C_(2)^T
  c.unfold_gram<float>(2, U1, d, Cf, d[2], KCf[2], d[2], 1.0, 0.0);
  c.unfold_gram<float>(2, U2, d, Cf, d[2], KCf[2] + (int64_t)d[2] * d[2], d[2], 1.0, 0.0);
  // KC_1: V1 = D x2 W2a + Ta x2 W2b ; V2 = Ta x2 W2a     (reuse U1/U2 storage after KC_2)
  float* V1 = U1; float* V2 = U2;
  c.mode_prod<float>(2, Wa[2], st_o[2], st_i[2], d[2], d[2], D, d, V1, 1.0, 0.0);
  c.mode_prod<float>(2, Wb[2], st_o[2], st_i[2], d[2], d[2], Ta, d, V1, 1.0, 1.0);
  c.mode_prod<float>(2, Wa[2], st_o[2], st_i[2], d[2], d[2], Ta, d, V2, 1.0, 0.0);
  c.unfold_gram<float>(1, V1, d, Cf, d[1], KCf[1], d[1], 1.0, 0.0);
  c.unfold_gram<float>(1, V2, d, Cf, d[1], KCf[1] + (int64_t)d[1] * d[1], d[1], 1.0, 0.0);
  // KC_0: Ea = Co x1 W1a, Eb = Co x1 W1b, F = Xo x1 W1a + Eb ; Z1 = F x2 W2a + Ea x2 W2b ; Z2 = Ea x2 W2a
  float* Ea = Ta; float* F = D;
  c.mode_prod<float>(1, Wa[1], st_o[1], st_i[1], d[1], d[1], Co, d, Ea, 1.0, 0.0);
  c.mode_prod<float>(1, Wa[1], st_o[1], st_i[1], d[1], d[1], Xo, d, F, 1.0, 0.0);
  c.mode_prod<float>(1, Wb[1], st_o[1], st_i[1], d[1], d[1], Co, d, F, 1.0, 1.0);
  float* Z1 = U1; float* Z2 = U2;
  c.mode_prod<float>(2, Wa[2], st_o[2], st_i[2], d[2], d[2], F, d, Z1, 1.0, 0.0);
  c.mode_prod<float>(2, Wb[2], st_o[2], st_i[2], d[2], d[2], Ea, d, Z1, 1.0, 1.0);
  c.mode_prod<float>(2, Wa[2], st_o[2], st_i[2], d[2], d[2], Ea, d, Z2, 1.0, 0.0);
  c.unfold_gram<float>(0, Z1, d, Cf, d[0], KCf[0], d[0], 1.0, 0.0);
  c.unfold_gram<float>(0, Z2, d, Cf, d[0], KCf[0] + (int64_t)d[0] * d[0], d[0], 1.0, 0.0);
  for (int i = 0; i < 3; ++i) c.to64(KCf[i], KC[i], 2 * (int64_t)d[i] * d[i]);
  if (sym) c.axpby(KC[1], KC[2], KC[1], 2 * (int64_t)d[1] * d[1], 1.0, nullptr, 1.0, nullptr);
  const int nm = sym ? 2 : 3;
  for (int i = 0; i < nm; ++i) {
    // K_i = beta * KC_i Ainv_i ;  L_i = -M_i K_i
    double* tk = c.p(c.L.tmpK[i]);
    c.matmul(KC[i], d[i], false, c.p(c.L.Ainv[i]), d[i], false, tk, d[i], 2 * d[i], d[i], d[i], 1.0, 0.0);
    c.axpby(tk, nullptr, Kout[i], 2 * (int64_t)d[i] * d[i], 1.0, hyper + 2, 0.0, nullptr);
    c.matmul(M[i], 2 * d[i], false, Kout[i], d[i], false, Lout[i], d[i], d[i], d[i], 2 * d[i], -1.0, 0.0);
  }
  if (sym && K_O && K_O != K_S) {
    c.axpby(K_S, nullptr, K_O, 2 * (int64_t)d[1] * d[1], 1.0, nullptr, 0.0, nullptr);
    c.axpby(L_S, nullptr, L_O, (int64_t)d[1] * d[1], 1.0, nullptr, 0.0, nullptr);
  }
  return finish(c, "rt_small_project");
}

extern "C" int rt_small_retract(const float* core, const float* dS_dir, const double* gram_R,
                                const double* gram_S, const double* gram_O, const double* hyper, int r0,
                                int r1, int r2, int sym, float* core_new, double* Z1_R, double* Z2_R,
                                double* Z1_S, double* Z2_S, double* Z1_O, double* Z2_O, double* Mn_R,
                                double* Mn_S, double* Mn_O, void* small_ws, void* stream) {
  RT_REQUIRE(small_ws != nullptr, "rt_small_retract: workspace is NULL");
  Ctx c{make_layout(r0, r1, r2, 0), (char*)small_ws, (cudaStream_t)stream};
  cudaStream_t s = c.s;
  const int d[3] = {r0, r1, r2};
  const double* gram[3] = {gram_R, gram_S, sym ? gram_S : gram_O};
  double* Z1[3] = {Z1_R, Z1_S, Z1_O};
  double* Z2[3] = {Z2_R, Z2_S, Z2_O};
  const double* lr = hyper + 0;
  double* C = c.p(c.L.C64);
  // C' = core - lr dS_dir
  double* Cp = c.T(0);
  {
    int blocks = (int)((c.L.c + 255) / 256); if (blocks > 1184) blocks = 1184;
    core_minus_lr_kernel<<<blocks, 256, 0, s>>>(core, dS_dir, lr, Cp, c.L.c); ++rt::g_launches;
  }
  // Gamma_i = lr^2 Gram_i = L_i L_i^T ; R_i = L_i^T
  SpdBatch b{};
  int nmax = 0;
  const int nm = sym ? 2 : 3;
  for (int i = 0; i < nm; ++i) {
    const int n2 = d[i] * d[i];
    scale_mat_kernel<<<cdiv(n2, 256), 256, 0, s>>>(gram[i], c.p(c.L.Gs[i]), n2, 1.0, lr, 1); ++rt::g_launches;
    b.p[i].G = c.p(c.L.Gs[i]); b.p[i].L = c.p(c.L.GL[i]); b.p[i].Linv = c.p(c.L.GLinv[i]);
    b.p[i].Ginv = nullptr; b.p[i].n = d[i];
    nmax = d[i] > nmax ? d[i] : nmax;
  }
  int rc = c.spd(b, nm, nmax);
  if (rc) return rc;
  const double* Lc[3] = {c.p(c.L.GL[0]), c.p(c.L.GL[1]), c.p(c.L.GL[sym ? 1 : 2])};
  const double* Linv[3] = {c.p(c.L.GLinv[0]), c.p(c.L.GLinv[1]), c.p(c.L.GLinv[sym ? 1 : 2])};
  // B_i = C x_i R_i,  R_i = L_i^T  => M[o,i] = L[i,o]: strides (sm_o, sm_i) = (1, d)
  double* Bk[3] = {c.T(1), c.T(2), c.T(3)};
  for (int i = 0; i < 3; ++i) c.mode_prod(i, Lc[i], 1, d[i], d[i], d[i], C, d, Bk[i], 1.0, 0.0);
  // unfolding Grams N_i (2r_i x 2r_i)
  double* Nn[3] = {c.p(c.L.Nn[0]), c.p(c.L.Nn[1]), c.p(c.L.Nn[2])};
  for (int i = 0; i < 3; ++i) {
    const int n = 2 * d[i];
    RT_CHECK_CUDA(cudaMemsetAsync(Nn[i], 0, sizeof(double) * n * n, s));
    c.unfold_gram(i, Cp, d, Cp, d[i], Nn[i], n, 1.0, 0.0);
    for (int j = 0; j < 3; ++j)
      if (j != i) c.unfold_gram(i, Bk[j], d, Bk[j], d[i], Nn[i], n, 1.0, 1.0);
    c.unfold_gram(i, Bk[i], d, Cp, d[i], Nn[i] + (int64_t)d[i] * n, n, 1.0, 0.0);
    c.unfold_gram(i, Bk[i], d, Bk[i], d[i], Nn[i] + (int64_t)d[i] * n + d[i], n, 1.0, 0.0);
    symmetrize_lower_kernel<<<cdiv(n * n, 256), 256, 0, s>>>(Nn[i], n); ++rt::g_launches;
  }
  if (sym) c.axpby(Nn[1], Nn[2], Nn[1], 4 * (int64_t)d[1] * d[1], 1.0, nullptr, 1.0, nullptr);
  if (c.err) return finish(c, "rt_small_retract");
  if (hosvd_route() == 1) {
    const double* Ain[3]; int nn[3]; int rr[3]; int ldy[3]; double* Yv[3]; void* sws[3];
    for (int i = 0; i < nm; ++i) {
      Ain[i] = Nn[i]; nn[i] = 2 * d[i]; rr[i] = d[i]; ldy[i] = 2 * d[i]; Yv[i] = c.p(c.L.Vf[i]);
      sws[i] = c.base + c.L.sub[i];
    }
    if ((rc = rt::subspace_batch(nm, Ain, nn, rr, Yv, ldy, sws, c.base + c.L.sub_shared, nullptr, s))) return rc;
  } else {
    const double* Ain[3]; int nn[3]; double* wv[3]; double* Vv[3]; void* ews[3];
    for (int i = 0; i < nm; ++i) {
      Ain[i] = Nn[i]; nn[i] = 2 * d[i]; wv[i] = c.p(c.L.Wv[i]); Vv[i] = c.p(c.L.Vf[i]);
      ews[i] = c.base + c.L.eig[i];
    }
    if ((rc = rt::eig_batch(nm, Ain, nn, wv, Vv, ews, s, hosvd_stop()))) return rc;   // see eig.cu
  }
  // Y_i = V_i[:, :r_i]  (2r_i x r_i, row stride 2r_i);  W_i = Y_i^T = [Y_ia^T | Y_ib^T]
  const double* Y[3] = {c.p(c.L.Vf[0]), c.p(c.L.Vf[1]), c.p(c.L.Vf[sym ? 1 : 2])};
  // core_new = T x_i Y_i^T through the block structure, in fp32 (the output is fp32): fp32 images of C', B_k
  // and of Y_i (dense [2 r_i, r_i]) live in the scratch temporaries T4..T7 (8 x c floats) and tmpM.
  float* CpF = c.Tf(4); float* BkF[3] = {c.Tf(4) + c.L.c, c.Tf(5), c.Tf(5) + c.L.c};
  c.to32plain(Cp, CpF, c.L.c);
  for (int i = 0; i < 3; ++i) c.to32plain(Bk[i], BkF[i], c.L.c);
  float* Yf[3];
  for (int i = 0; i < 3; ++i) Yf[i] = reinterpret_cast<float*>(c.p(c.L.tmpM[i]));
  for (int i = 0; i < nm; ++i) {
    const int total = 2 * d[i] * d[i];
    f64_to_f32_strided_kernel<<<cdiv(total, 256), 256, 0, s>>>(Y[i], 2 * d[i], Yf[i], d[i], 2 * d[i], d[i]);
    ++rt::g_launches;
  }
  if (sym) Yf[2] = Yf[1];
  const float* Wa[3]; const float* Wb[3]; int64_t st_o[3], st_i[3];
  for (int i = 0; i < 3; ++i) {
    Wa[i] = Yf[i];                              // W_ia[o, k] = Y[k, o]      -> strides (1, r_i)
    Wb[i] = Yf[i] + (int64_t)d[i] * d[i];       // W_ib[o, k] = Y[r_i + k, o]
    st_o[i] = 1; st_i[i] = d[i];
  }
  const float* Cb[3] = {BkF[0], BkF[1], BkF[2]};
  float* Dn = c.Tf(6); float* E1 = c.Tf(6) + c.L.c; float* E2 = c.Tf(7); float* U1n = c.Tf(7) + c.L.c;
  // U2 reuses E1 (dead once U1 exists), the result reuses D
  grouped_contract<float>(c, CpF, Cb, Wa, Wb, st_o, st_i, Dn, E1, E2, U1n, E1, Dn);
  RT_CHECK_CUDA(cudaMemcpyAsync(core_new, Dn, sizeof(float) * c.L.c, cudaMemcpyDeviceToDevice, s));
  // Z1_i = Y_ia ; Z2_i = -lr * L_i^-T Y_ib
  for (int i = 0; i < nm; ++i) {
    const int64_t n = 2 * d[i];
    c.matmul(Linv[i], d[i], true, Y[i] + (int64_t)d[i] * n, n, false, c.p(c.L.tmpK[i]), d[i], d[i], d[i], d[i], -1.0, 0.0);
    c.axpby(c.p(c.L.tmpK[i]), nullptr, Z2[i], (int64_t)d[i] * d[i], 1.0, lr, 0.0, nullptr);
    RT_CHECK_CUDA(cudaMemcpy2DAsync(Z1[i], sizeof(double) * d[i], Y[i], sizeof(double) * n,
                                    sizeof(double) * d[i], d[i], cudaMemcpyDeviceToDevice, s));
  }
  // Transport Grams of the NEXT fit() without touching N-sized data: with U^T U = I and
  // U^T dV = 0,  U_new^T [U | dV] = [Z1^T | Z2^T (dV^T dV)]   (r_i x 2 r_i)
  double* Mn[3] = {Mn_R, Mn_S, Mn_O};
  for (int i = 0; i < nm; ++i) {
    if (!Mn[i]) continue;
    const int64_t n2 = 2 * d[i];
    transpose_into_kernel<<<cdiv(d[i] * d[i], 256), 256, 0, s>>>(Z1[i], d[i], Mn[i], n2);
    ++rt::g_launches;
    c.matmul(Z2[i], d[i], true, gram[i], d[i], false, Mn[i] + d[i], n2, d[i], d[i], d[i], 1.0, 0.0);
  }
  if (sym && Z1_O && Z1_O != Z1_S) {
    c.axpby(Z1_S, nullptr, Z1_O, (int64_t)d[1] * d[1], 1.0, nullptr, 0.0, nullptr);
    c.axpby(Z2_S, nullptr, Z2_O, (int64_t)d[1] * d[1], 1.0, nullptr, 0.0, nullptr);
  }
  return finish(c, "rt_small_retract");
}
