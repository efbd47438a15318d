// kernels_leaf.cu — leaf boundaries, the fused per-leaf fit + forward/error pass, and the
// summary statistics.
//
// Reference control flow being replaced (rmi_lib/src/train/two_layer.rs):
//   :131-145  split_idx = first key whose clamped top prediction reaches N/2
//   :147-175  build_models_from over [0, split) and [split+1, n)  (the key AT split_idx is in
//             no leaf's training set)
//   :20-99    per leaf: training vector = [last key of previous non-empty leaf] + own keys +
//             [first key of next non-empty leaf] (neither across the half boundary), model
//             fitted by train_model(layer2, vector)
//   :178-197  LowerBoundCorrection::new (lower_bound_correction.rs:91-137) and constant models
//             for empty leaves
//   :207-217  forward pass over every key: per-leaf (count, max |pred - offset|)
//   :226-259  widening by the neighbours' keys and the longest duplicate run
//
// B200 formulation.  The clamped top prediction is non-decreasing over the sorted keys (the
// reference asserts it, :50), so leaf j owns the contiguous index range [S[j], S[j+1]) with
// S[j] = first index whose prediction is >= j.  One streaming pass produces S (k_bounds);
// after that every quantity the reference derives by walking all n keys three more times is a
// function of the keys in [S[j]-1, S[j+1]] alone, so ONE kernel (k_leaf) fits leaf j, replaces
// it by a constant if it is empty, evaluates it on its own keys and widens the bound — the
// leaf's keys are touched by one lane while they are hot in L1/L2.  The fit is the reference's
// order-dependent recurrence run in the reference's order by a single lane per leaf, hence
// bit-identical; parallelism comes from the N independent leaves.
#include "device_util.cuh"
#include "kernels.h"

namespace rmi {

namespace {

constexpr int BOUNDS_THREADS = 256;
constexpr int LEAF_THREADS = 128;
constexpr int RCP_TABLE = 1024;

__device__ __forceinline__ void set_status(BuildAux* aux, unsigned bit) { atomicOr(&aux->status, bit); }

__host__ __device__ constexpr bool top_needs_bounds_check(int kind) {
  // cubic_spline.rs:184, radix.rs:75,164, balanced_radix.rs:164, histogram.rs:103
  return !(kind == M_CUBIC || kind == M_RADIX || kind == M_RADIX_TABLE || kind == M_BRADIX || kind == M_HISTOGRAM);
}

__global__ void k_fill(u64* __restrict__ p, u64 len, u64 v) {
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) p[i] = v;
}

// S[j] = first index i with min(N-1, top(key_i)) >= j.  S is pre-filled with n.
template <class T, int TOP>
__global__ void __launch_bounds__(BOUNDS_THREADS)
k_bounds(const T* __restrict__ keys, u64 n, const TopModel* __restrict__ top_ptr, u64 N, u64* __restrict__ S,
         BuildAux* aux) {
  TopModel m = *top_ptr;
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T k = keys[i];
    u64 p = top_predict<TOP>(m, k);
    if (!top_needs_bounds_check(TOP) && p >= N) set_status(aux, ST_TOP_OUT_OF_BOUNDS);
    u64 t = p < N - 1 ? p : N - 1;
    if (i == 0) {
      for (u64 q = 0; q <= t; ++q) S[q] = 0;
    } else {
      T kp = keys[i - 1];
      if (k < kp) set_status(aux, ST_NOT_SORTED);
      u64 pp = top_predict<TOP>(m, kp);
      u64 tp = pp < N - 1 ? pp : N - 1;
      if (t < tp) set_status(aux, ST_NON_MONOTONE);
      for (u64 q = tp + 1; q <= t; ++q) S[q] = i;
    }
  }
}

// two_layer.rs:131-159
template <class T, int TOP>
__global__ void k_split(const T* __restrict__ keys, u64 n, const TopModel* __restrict__ top_ptr, u64 N,
                        const u64* __restrict__ S, BuildAux* aux) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  u64 split = S[N / 2];
  aux->split_idx = split;
  if (split >= n) { aux->has_split = 0; aux->split_target = 0; return; }
  aux->has_split = 1;
  if (split == 0) set_status(aux, ST_SPLIT_AT_ZERO);
  if (split + 1 >= n) set_status(aux, ST_SPLIT_AT_END);
  TopModel m = *top_ptr;
  u64 p = top_predict<TOP>(m, keys[split]);
  aux->split_target = p < N - 1 ? p : N - 1;
}

// ------------------------------------------------------------------------------------------
// Per-leaf training stream: (key, offset) items in the order train_model sees them.
// ------------------------------------------------------------------------------------------
template <class T> struct LeafRange {
  const T* __restrict__ keys;
  u64 n;
  u64 lo, hi;          // all keys of the leaf: [S[j], S[j+1])
  u64 own_lo, own_hi;  // keys in the leaf's training set (the split key excluded)
  u64 half_lo;         // first index of the leaf's half
  int mode;            // 0: empty vector, 1: [P] own [Nx], 2: single item keys[half_lo]
  bool has_prev, has_next;
  __device__ __forceinline__ u64 vec_len() const {
    if (mode == 0) return 0;
    if (mode == 2) return 1;
    return (own_hi - own_lo) + (has_prev ? 1 : 0) + (has_next ? 1 : 0);
  }
};

// Calls fn(key, offset) for every item of the vector (NOT including the drained-iterator
// repeat).  Offsets are the duplicate-fixed global offsets F.
template <class T, class Fn> __device__ __forceinline__ void walk_vector(const LeafRange<T>& r, Fn&& fn) {
  if (r.mode == 0) return;
  if (r.mode == 2) { fn(r.keys[r.half_lo], run_start(r.keys, r.half_lo)); return; }
  T pk = T();
  u64 pF = 0;
  bool have = false;
  if (r.own_lo > 0) { pk = r.keys[r.own_lo - 1]; pF = run_start(r.keys, r.own_lo - 1); have = true; }
  if (r.has_prev) fn(pk, pF);
  for (u64 i = r.own_lo; i < r.own_hi; ++i) {
    T k = r.keys[i];
    u64 F = (have && k == pk) ? pF : i;
    fn(k, F);
    pk = k; pF = F; have = true;
  }
  if (r.has_next) fn(r.keys[r.own_hi], r.own_hi);
}
// First / last raw item of the vector (RMITrainingData::get, models/mod.rs:268-270).
template <class T> __device__ __forceinline__ void vector_first(const LeafRange<T>& r, T& k, u64& y) {
  u64 i = r.mode == 2 ? r.half_lo : (r.has_prev ? r.own_lo - 1 : r.own_lo);
  k = r.keys[i]; y = run_start(r.keys, i);
}
template <class T> __device__ __forceinline__ void vector_last(const LeafRange<T>& r, T& k, u64& y) {
  u64 i = r.mode == 2 ? r.half_lo : (r.has_next ? r.own_hi : r.own_hi - 1);
  k = r.keys[i]; y = run_start(r.keys, i);
}
// Raw item at vector position p (0-based).
template <class T> __device__ __forceinline__ void vector_at(const LeafRange<T>& r, u64 p, T& k, u64& y) {
  u64 i = r.mode == 2 ? r.half_lo : (r.has_prev ? r.own_lo - 1 + p : r.own_lo + p);
  k = r.keys[i]; y = run_start(r.keys, i);
}

// The reference's Welford step (linear.rs:24-34) with the two count divisions done by
// div_by_count (rust_math.cuh): bit-identical to IEEE division, 3 FP64 ops instead of ~20.
struct LeafWelford {
  double mean_x, mean_y, c, m2, nf;
  const double* rcp;
  __device__ __forceinline__ void init(const double* table) { mean_x = mean_y = c = m2 = nf = 0.0; rcp = table; }
  __device__ __forceinline__ void push(double x, double y) {
    nf = __dadd_rn(nf, 1.0);
    double rc = nf < (double)RCP_TABLE ? rcp[(int)nf] : __drcp_rn(nf);
    double dx = __dadd_rn(x, -mean_x);
    mean_x = __dadd_rn(mean_x, div_by_count(dx, nf, rc));
    mean_y = __dadd_rn(mean_y, div_by_count(__dadd_rn(y, -mean_y), nf, rc));
    c = __dadd_rn(c, __dmul_rn(dx, __dadd_rn(y, -mean_y)));
    double dx2 = __dadd_rn(x, -mean_x);
    m2 = __dadd_rn(m2, __dmul_rn(dx, dx2));
  }
  __device__ __forceinline__ bool finish(double& alpha, double& beta) const {   // linear.rs:36-58
    if (nf == 0.0) { alpha = 0.0; beta = 0.0; return true; }
    if (nf == 1.0) { alpha = mean_y; beta = 0.0; return true; }
    double nm1 = __dadd_rn(nf, -1.0);
    double cov = __ddiv_rn(c, nm1), var = __ddiv_rn(m2, nm1);
    if (!(var >= 0.0)) { alpha = 0.0; beta = 0.0; return false; }
    if (var == 0.0) { alpha = mean_y; beta = 0.0; return true; }
    beta = __ddiv_rn(cov, var);
    alpha = __dadd_rn(mean_y, -__dmul_rn(beta, mean_x));
    return true;
  }
};

__device__ __forceinline__ double scale3(double v, double mn, double mx) {
  return __ddiv_rn(__dadd_rn(v, -mn), __dadd_rn(mx, -mn));
}

// train_model(layer2, vector) for every leaf model type.  f receives Model::params().
template <class T, int LEAF>
__device__ __forceinline__ void fit_leaf(const LeafRange<T>& r, const double* rcp, double* f, BuildAux* aux) {
  const u64 L = r.vec_len();
  if (LEAF == M_LINEAR || LEAF == M_LOGLINEAR) {
    // linear.rs:79-83 / :61-72,169-173 — drained stream: vector + repeat of the final item
    LeafWelford w;
    w.init(rcp);
    T lk = T(); u64 ly = 0;
    walk_vector(r, [&](T k, u64 y) {
      lk = k; ly = y;
      double yy = __ull2double_rn(y);
      if (LEAF == M_LOGLINEAR) { yy = log(yy); if (!isfinite(yy)) return; }
      w.push(Key<T>::as_float(k), yy);
    });
    if (L > 0) {
      double yy = __ull2double_rn(ly);
      if (LEAF == M_LOGLINEAR) yy = log(yy);
      if (LEAF == M_LINEAR || isfinite(yy)) w.push(Key<T>::as_float(lk), yy);
    }
    if (!w.finish(f[0], f[1])) set_status(aux, ST_NEG_VARIANCE);
  } else if (LEAF == M_ROBUST_LINEAR) {
    // linear.rs:239-260 — skip(bnd).take(len - 2*bnd): never drains the iterator
    if (L == 0) { f[0] = 0.0; f[1] = 0.0; return; }
    u64 bnd = f64_to_u64_sat(__dmul_rn(__ull2double_rn(L), 0.0001));
    if (bnd < 1) bnd = 1;
    if (!(bnd * 2 + 1 < L)) { set_status(aux, ST_ROBUST_TOO_SMALL); f[0] = 0.0; f[1] = 0.0; return; }
    LeafWelford w;
    w.init(rcp);
    u64 pos = 0;
    walk_vector(r, [&](T k, u64 y) {
      if (pos >= bnd && pos < L - bnd) w.push(Key<T>::as_float(k), __ull2double_rn(y));
      ++pos;
    });
    if (!w.finish(f[0], f[1])) set_status(aux, ST_NEG_VARIANCE);
  } else if (LEAF == M_LINEAR_SPLINE || LEAF == M_CUBIC) {
    // linear_spline.rs:13-35
    double la, lb;
    T k0 = T(), k1 = T(); u64 y0 = 0, y1 = 0;
    if (L > 0) { vector_first(r, k0, y0); vector_last(r, k1, y1); }
    if (L == 0) { la = 0.0; lb = 0.0; }
    else if (L == 1 || k0 == k1) { la = __ull2double_rn(y0); lb = 0.0; }
    else {
      double x0 = Key<T>::as_float(k0), x1 = Key<T>::as_float(k1);
      double slope = __ddiv_rn(__dadd_rn(__ull2double_rn(y0), -__ull2double_rn(y1)), __dadd_rn(x0, -x1));
      la = __dadd_rn(__ull2double_rn(y0), -__dmul_rn(slope, x0));
      lb = slope;
    }
    if (LEAF == M_LINEAR_SPLINE) { f[0] = la; f[1] = lb; return; }
    // cubic_spline.rs:18-101
    double a, b, c, d;
    if (L == 0) { a = 0.0; b = 0.0; c = 1.0; d = 0.0; }
    else {
      bool uniq = false;
      if (L > 1) walk_vector(r, [&](T k, u64) { if (k != k0) uniq = true; });
      if (L == 1 || !uniq) { a = b = c = 0.0; d = __ull2double_rn(y0); }
      else {
        double xmin = Key<T>::as_float(k0), ymin = __ull2double_rn(y0);
        double xmax = Key<T>::as_float(k1), ymax = __ull2double_rn(y1);
        bool found1 = false; double sxn = 0.0, syn = 0.0;
        walk_vector(r, [&](T k, u64 y) {
          if (found1) return;
          double sx = scale3(Key<T>::as_float(k), xmin, xmax);
          if (sx > 0.0) { found1 = true; sxn = sx; syn = scale3(__ull2double_rn(y), ymin, ymax); }
        });
        bool found2 = false; double sxp = 0.0, syp = 0.0;
        for (u64 p = L; p-- > 0;) {
          T k; u64 y;
          vector_at(r, p, k, y);
          double sx = scale3(Key<T>::as_float(k), xmin, xmax);
          if (sx < 1.0) { found2 = true; sxp = sx; syp = scale3(__ull2double_rn(y), ymin, ymax); break; }
        }
        if (!found1 || !found2) { set_status(aux, ST_CUBIC_UNWRAP); a = b = c = d = 0.0; }
        else {
          double m1 = __ddiv_rn(syn, sxn);
          double m2 = __ddiv_rn(__dadd_rn(1.0, -syp), __dadd_rn(1.0, -sxp));
          double ss = __dadd_rn(__dmul_rn(m1, m1), __dmul_rn(m2, m2));
          if (ss > 9.0) {
            double tau = __ddiv_rn(3.0, __dsqrt_rn(ss));
            m1 = __dmul_rn(m1, tau);
            m2 = __dmul_rn(m2, tau);
          }
          double d3 = cube_dd(__dadd_rn(xmax, -xmin));
          a = __ddiv_rn(__dadd_rn(__dadd_rn(m1, m2), -2.0), d3);
          double t1 = __dmul_rn(xmax, __dadd_rn(__dadd_rn(__dmul_rn(2.0, m1), m2), -3.0));
          double t2 = __dmul_rn(xmin, __dadd_rn(__dadd_rn(m1, __dmul_rn(2.0, m2)), -3.0));
          b = __ddiv_rn(-__dadd_rn(t1, t2), d3);
          double xmax2 = __dmul_rn(xmax, xmax), xmin2 = __dmul_rn(xmin, xmin);
          double u1 = __dmul_rn(m1, xmax2), u2 = __dmul_rn(m2, xmin2);
          double u3 = __dmul_rn(__dmul_rn(xmax, xmin),
                                __dadd_rn(__dadd_rn(__dmul_rn(2.0, m1), __dmul_rn(2.0, m2)), -6.0));
          c = __ddiv_rn(__dadd_rn(__dadd_rn(u1, u2), u3), d3);
          double v2 = __dmul_rn(__dmul_rn(xmax, xmin), __dadd_rn(m2, -3.0));
          d = __ddiv_rn(__dmul_rn(-xmin, __dadd_rn(__dadd_rn(u1, v2), xmin2)), d3);
          double dy = __dadd_rn(ymax, -ymin);
          a = __dmul_rn(a, dy); b = __dmul_rn(b, dy); c = __dmul_rn(c, dy); d = __dmul_rn(d, dy);
          d = __dadd_rn(d, ymin);
        }
      }
    }
    // cubic_spline.rs:113-135: keep the linear spline if its L1 error is strictly lower
    double cf[4] = {a, b, c, d}, lf[2] = {la, lb};
    double our_error = 0.0, lin_error = 0.0;
    T lk = T(); u64 ly = 0;
    auto acc = [&](T k, u64 y) {
      double x = Key<T>::as_float(k), yy = __ull2double_rn(y);
      our_error = __dadd_rn(our_error, fabs(__dadd_rn(predict_float<M_CUBIC>(cf, x), -yy)));
      lin_error = __dadd_rn(lin_error, fabs(__dadd_rn(predict_float<M_LINEAR>(lf, x), -yy)));
      lk = k; ly = y;
    };
    walk_vector(r, acc);
    if (L > 0) acc(lk, ly);
    if (lin_error < our_error) { f[0] = 0.0; f[1] = 0.0; f[2] = lb; f[3] = la; }
    else { f[0] = a; f[1] = b; f[2] = c; f[3] = d; }
  } else {  // M_NORMAL / M_LOGNORMAL — normal.rs:28-76
    double scale = -INFINITY, mean = 0.0, stdev = 0.0;
    double nf = __ull2double_rn(L);
    T lk = T(); u64 ly = 0;
    auto tx = [&](T k) {
      double x = Key<T>::as_float(k);
      if (LEAF == M_LOGNORMAL) { double l = log(x); x = isfinite(l) ? l : 0.0; }
      return x;
    };
    auto p1 = [&](T k, u64 y) {
      mean = __dadd_rn(mean, __ddiv_rn(tx(k), nf));
      scale = rust_fmax(scale, __ull2double_rn(y));
      lk = k; ly = y;
    };
    walk_vector(r, p1);
    if (L > 0) p1(lk, ly);
    auto p2 = [&](T k, u64) { double dlt = __dadd_rn(tx(k), -mean); stdev = __dadd_rn(stdev, __dmul_rn(dlt, dlt)); };
    walk_vector(r, p2);
    if (L > 0) p2(lk, ly);
    stdev = __dsqrt_rn(__ddiv_rn(stdev, nf));
    f[0] = mean; f[1] = stdev; f[2] = scale;
  }
}

// set_to_constant_model (linear.rs:116-119,293-296, linear_spline.rs:79-82,
// cubic_spline.rs:188-191, default models/mod.rs:761-763)
template <int LEAF> __device__ __forceinline__ bool set_constant(double* f, u64 c) {
  if (LEAF == M_LINEAR || LEAF == M_ROBUST_LINEAR || LEAF == M_LINEAR_SPLINE) {
    f[0] = __ull2double_rn(c); f[1] = 0.0; return true;
  } else if (LEAF == M_CUBIC) {
    f[0] = 0.0; f[1] = 0.0; f[2] = 0.0; f[3] = __ull2double_rn(c); return true;
  }
  return false;
}

template <class T, int LEAF>
__global__ void __launch_bounds__(LEAF_THREADS)
k_leaf(const T* __restrict__ keys, u64 n, u64 N, const u64* __restrict__ S, BuildAux* aux,
       double* __restrict__ params, u64* __restrict__ errors, u64* __restrict__ counts) {
  __shared__ double s_rcp[RCP_TABLE];
  for (int c = threadIdx.x; c < RCP_TABLE; c += blockDim.x) s_rcp[c] = c ? __drcp_rn((double)c) : 0.0;
  __syncthreads();
  u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  constexpr int PPM = leaf_params_per_model(LEAF);

  LeafRange<T> r;
  r.keys = keys; r.n = n;
  r.lo = S[j]; r.hi = S[j + 1];
  // which half does leaf j belong to (two_layer.rs:147-175)
  u64 half_hi, first_leaf;
  if (aux->has_split) {
    u64 split = aux->split_idx, st = aux->split_target;
    if (j < st) { r.half_lo = 0; half_hi = split; first_leaf = 0; }
    else { r.half_lo = split + 1; half_hi = n; first_leaf = st; }
  } else { r.half_lo = 0; half_hi = n; first_leaf = 0; }
  r.own_lo = r.lo > r.half_lo ? r.lo : r.half_lo;
  r.own_hi = r.hi < half_hi ? r.hi : half_hi;
  if (r.own_hi > r.own_lo) {
    r.mode = 1;
    r.has_prev = r.own_lo > r.half_lo;
    r.has_next = r.own_hi < half_hi;
  } else {
    r.own_hi = r.own_lo;
    r.has_prev = r.has_next = false;
    // the half's first leaf, if it owns no key, is trained on the half's first item alone
    // (two_layer.rs:52-62 with an empty second_layer_data); other empty leaves on empty data.
    r.mode = (j == first_leaf && r.half_lo < half_hi) ? 2 : 0;
  }

  double f[4] = {0.0, 0.0, 0.0, 0.0};
  fit_leaf<T, LEAF>(r, s_rcp, f, aux);

  // two_layer.rs:186-197: empty leaves (lower-bound-correction sense) except the last
  const u64 next_idx = S[j + 1];                                  // lb.next_index(j)
  if (j + 1 < N && r.lo == r.hi) {
    if (!set_constant<LEAF>(f, next_idx)) atomicAdd(&aux->could_not_replace, 1ull);
  }

  // two_layer.rs:207-217 forward pass over the leaf's own keys + longest run
  // (lower_bound_correction.rs:101-119: a run is recorded when the NEXT run starts, so the
  // data set's final run never is)
  u64 max_err = 0, run_max = 0;
  {
    T pk = T();
    u64 F = r.lo, run = 0;
    for (u64 i = r.lo; i < r.hi; ++i) {
      T k = keys[i];
      if (i == r.lo || k != pk) { if (run > run_max) run_max = run; run = 1; F = i; pk = k; }
      else run += 1;
      u64 pred = predict_int_f<LEAF>(f, Key<T>::as_float(k));
      u64 e = error_between(pred, F, n);
      if (e > max_err) max_err = e;
    }
    if (r.hi < n && run > run_max) run_max = run;
  }
  u64 cnt = r.hi - r.lo;
  if (r.hi == n && r.lo < r.hi) cnt += 1;   // the drained iterator's repeated final item

  // two_layer.rs:226-259 widening
  T next_key = next_idx < n ? keys[next_idx] : Key<T>::max_value();
  T prev_key = r.lo > 0 ? keys[r.lo - 1] : Key<T>::zero_value();
  u64 first_idx = j == 0 ? S[1] : r.lo;                            // lb.next_index(max(j-1, 0))
  u64 up = predict_int_f<LEAF>(f, Key<T>::as_float(Key<T>::minus_epsilon(next_key)));
  u64 upper_error = error_between(up, next_idx + 1, n);
  u64 lp = predict_int_f<LEAF>(f, Key<T>::as_float(Key<T>::plus_epsilon(prev_key)));
  u64 lower_error = error_between(lp, first_idx, n);
  u64 new_err = max_err;
  if (upper_error > new_err) new_err = upper_error;
  if (lower_error > new_err) new_err = lower_error;
  new_err += run_max;

#pragma unroll
  for (int q = 0; q < PPM; ++q) params[j * PPM + q] = f[q];
  errors[j] = new_err;
  counts[j] = cnt;
}

// ------------------------------------------------------------------------------------------
// Summary statistics (two_layer.rs:267-284): fixed-shape tree reductions over the N leaves.
// partial per block: {max_err, max_idx, sum_n_err (bits), sum_l2, sum_log2}
// ------------------------------------------------------------------------------------------
constexpr int STATS_THREADS = 256;
constexpr int STATS_MAX_BLOCKS = 1024;

struct StatsPartial { u64 max_err, max_idx, sum_ne; double l2, lg; };

__device__ __forceinline__ void stats_merge(u64& me, u64& mi, u64 oe, u64 oi) {
  // max_by_key keeps the LAST maximum: larger error wins, ties go to the larger index
  if (oe > me || (oe == me && oi > mi)) { me = oe; mi = oi; }
}

__global__ void __launch_bounds__(STATS_THREADS)
k_stats_partial(u64 n, u64 N, const u64* __restrict__ errors, const u64* __restrict__ counts,
                StatsPartial* __restrict__ out) {
  __shared__ double smd[32];
  __shared__ u64 smu[32];
  __shared__ u64 sme[32], smi[32];
  double nf = __ull2double_rn(n);
  u64 me = 0, mi = 0, sne = 0;
  double l2 = 0.0, lg = 0.0;
  u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    u64 e = errors[j], c = counts[j];
    stats_merge(me, mi, e, j);
    u64 ne = c * e;
    sne += ne;
    double v = __ull2double_rn(ne);
    l2 += __ddiv_rn(__dmul_rn(v, v), nf);
    lg += __dmul_rn(__ull2double_rn(c), log2(__ull2double_rn(2ull * e + 2ull)));
  }
  // max reduction
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    u64 oe = __shfl_down_sync(0xffffffffu, me, o), oi = __shfl_down_sync(0xffffffffu, mi, o);
    stats_merge(me, mi, oe, oi);
  }
  if (lane == 0) { sme[w] = me; smi[w] = mi; }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    me = lane < nw ? sme[lane] : 0; mi = lane < nw ? smi[lane] : 0;
    for (int o = 16; o > 0; o >>= 1) {
      u64 oe = __shfl_down_sync(0xffffffffu, me, o), oi = __shfl_down_sync(0xffffffffu, mi, o);
      stats_merge(me, mi, oe, oi);
    }
  }
  u64 r_ne = block_sum_u64(sne, smu);
  double r_l2 = block_sum(l2, smd), r_lg = block_sum(lg, smd);
  if (threadIdx.x == 0) {
    StatsPartial p;
    p.max_err = me; p.max_idx = mi; p.sum_ne = r_ne; p.l2 = r_l2; p.lg = r_lg;
    out[blockIdx.x] = p;
  }
}
__global__ void k_stats_finish(const StatsPartial* __restrict__ parts, int nblocks, BuildAux* aux) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  u64 me = 0, mi = 0, sne = 0;
  double l2 = 0.0, lg = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    stats_merge(me, mi, parts[b].max_err, parts[b].max_idx);
    sne += parts[b].sum_ne;
    l2 += parts[b].l2;
    lg += parts[b].lg;
  }
  aux->max_error = me; aux->max_error_idx = mi; aux->sum_n_err = sne; aux->sum_l2 = l2; aux->sum_log2 = lg;
}

int grid_cap(u64 n, int threads, int cap) {
  u64 blocks = (n + threads - 1) / threads;
  if (blocks > (u64)cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <class T, int TOP>
void launch_bounds(const Launch& L, const T* keys, u64 n, const TopModel* d_top, u64 N, u64* d_S, BuildAux* d_aux) {
  k_fill<<<grid_cap(N + 1, BOUNDS_THREADS, L.num_sms * 8), BOUNDS_THREADS, 0, L.stream>>>(d_S, N + 1, n);
  count_launch();
  k_bounds<T, TOP><<<grid_cap(n, BOUNDS_THREADS, L.num_sms * 8), BOUNDS_THREADS, 0, L.stream>>>(keys, n, d_top, N, d_S, d_aux);
  count_launch();
  k_split<T, TOP><<<1, 32, 0, L.stream>>>(keys, n, d_top, N, d_S, d_aux);
  count_launch();
}

template <class T, int LEAF>
void launch_leaf(const Launch& L, const T* keys, u64 n, u64 N, const u64* d_S, BuildAux* d_aux, double* d_params,
                 u64* d_errors, u64* d_counts) {
  u64 blocks = (N + LEAF_THREADS - 1) / LEAF_THREADS;
  k_leaf<T, LEAF><<<(unsigned)blocks, LEAF_THREADS, 0, L.stream>>>(keys, n, N, d_S, d_aux, d_params, d_errors, d_counts);
  count_launch();
}

}  // namespace

template <class T>
void compute_leaf_bounds(const Launch& L, const T* keys, u64 n, int top_kind, const TopModel* d_top, u64 N, u64* d_S,
                         BuildAux* d_aux) {
  switch (top_kind) {
    case M_LINEAR:
    case M_ROBUST_LINEAR:
    case M_LINEAR_SPLINE: launch_bounds<T, M_LINEAR>(L, keys, n, d_top, N, d_S, d_aux); break;
    case M_CUBIC: launch_bounds<T, M_CUBIC>(L, keys, n, d_top, N, d_S, d_aux); break;
    case M_LOGLINEAR: launch_bounds<T, M_LOGLINEAR>(L, keys, n, d_top, N, d_S, d_aux); break;
    case M_NORMAL: launch_bounds<T, M_NORMAL>(L, keys, n, d_top, N, d_S, d_aux); break;
    case M_LOGNORMAL: launch_bounds<T, M_LOGNORMAL>(L, keys, n, d_top, N, d_S, d_aux); break;
    case M_RADIX: launch_bounds<T, M_RADIX>(L, keys, n, d_top, N, d_S, d_aux); break;
    case M_RADIX_TABLE: launch_bounds<T, M_RADIX_TABLE>(L, keys, n, d_top, N, d_S, d_aux); break;
    case M_BRADIX: launch_bounds<T, M_BRADIX>(L, keys, n, d_top, N, d_S, d_aux); break;
    case M_HISTOGRAM: launch_bounds<T, M_HISTOGRAM>(L, keys, n, d_top, N, d_S, d_aux); break;
    default: break;
  }
}

template <class T>
void fit_leaves(const Launch& L, const T* keys, u64 n, int leaf_kind, u64 N, const u64* d_S, BuildAux* d_aux,
                double* d_params, u64* d_errors, u64* d_counts) {
  switch (leaf_kind) {
    case M_LINEAR: launch_leaf<T, M_LINEAR>(L, keys, n, N, d_S, d_aux, d_params, d_errors, d_counts); break;
    case M_ROBUST_LINEAR: launch_leaf<T, M_ROBUST_LINEAR>(L, keys, n, N, d_S, d_aux, d_params, d_errors, d_counts); break;
    case M_LINEAR_SPLINE: launch_leaf<T, M_LINEAR_SPLINE>(L, keys, n, N, d_S, d_aux, d_params, d_errors, d_counts); break;
    case M_CUBIC: launch_leaf<T, M_CUBIC>(L, keys, n, N, d_S, d_aux, d_params, d_errors, d_counts); break;
    case M_LOGLINEAR: launch_leaf<T, M_LOGLINEAR>(L, keys, n, N, d_S, d_aux, d_params, d_errors, d_counts); break;
    case M_NORMAL: launch_leaf<T, M_NORMAL>(L, keys, n, N, d_S, d_aux, d_params, d_errors, d_counts); break;
    case M_LOGNORMAL: launch_leaf<T, M_LOGNORMAL>(L, keys, n, N, d_S, d_aux, d_params, d_errors, d_counts); break;
    default: break;
  }
}

size_t stats_scratch_bytes(u64) { return sizeof(StatsPartial) * STATS_MAX_BLOCKS; }

void leaf_statistics(const Launch& L, u64 n, u64 N, const u64* d_errors, const u64* d_counts, BuildAux* d_aux,
                     void* scratch) {
  int g = grid_cap(N, STATS_THREADS, STATS_MAX_BLOCKS);
  k_stats_partial<<<g, STATS_THREADS, 0, L.stream>>>(n, N, d_errors, d_counts, (StatsPartial*)scratch);
  count_launch();
  k_stats_finish<<<1, 32, 0, L.stream>>>((const StatsPartial*)scratch, g, d_aux);
  count_launch();
}

#define INST(T)                                                                                                  \
  template void compute_leaf_bounds<T>(const Launch&, const T*, u64, int, const TopModel*, u64, u64*, BuildAux*); \
  template void fit_leaves<T>(const Launch&, const T*, u64, int, u64, const u64*, BuildAux*, double*, u64*, u64*);
INST(u64)
INST(u32)
INST(double)
#undef INST

}  // namespace rmi
