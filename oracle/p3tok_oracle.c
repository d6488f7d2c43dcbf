/*
 * p3tok_oracle.c - CPU restatement of the reference point-patch tokenizer's INDEX work.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this library; the product path (p3tok/) never does.
 *
 * Parity status: the reference ships no golden vectors or known-answer tests for this
 * path, so this restatement is pinned against outputs of the reference
 * itself, generated in the build container by tests/golden/make_golden.py and committed
 * under tests/golden/ (see tests/test_oracle_vs_golden.py).
 *
 * The arithmetic lives in third-party torch (pinned torch==2.7.1+cu128 by the reference's
 * requirements.txt:44; torch 2.11.0 + MKL 2024.2 in this image).  What is restated here is
 * the exact fp32 operation order those calls perform, verified bit-for-bit in the build
 * container:
 *   FPS   dist = ((dx*dx) + (dy*dy)) + (dz*dz), every op individually rounded
 *         (torch.sum((xyz - c) ** 2, -1);  src/data/sampler.py:26, src/models/pix4point.py:44)
 *   APF   d = ((-2*dot) + |c|^2) + |p|^2, dot = fma(cz,pz, fma(cy,py, cx*px)),
 *         |v|^2 = ((vx*vx)+(vy*vy))+(vz*vz)             (src/data/sampler.py:59-61)
 *   P4P   t = fma(1,|p|^2, fma(|c|^2,1, fma(-2cz,pz, fma(-2cy,py, (-2cx)*px))));
 *         d = sqrt(max(t,0))    (torch.cdist mm path, src/models/pix4point.py:87)
 *         NOTE: torch's CPU sqrt kernel is not correctly rounded on this host (0.6 % of
 *         values are 1 ulp low); this oracle uses IEEE sqrtf, which is what the CUDA
 *         device computes, and tests allow a 1-ulp tie window against the torch CPU run.
 * Ties: FPS argmax keeps the LOWEST index (torch.max, sampler.py:28).  kNN selection and
 * order are canonical here: ascending (distance, index).  torch.topk's tie order is
 * libstdc++-defined, so parity with the reference is "equal up to exact ties".
 *
 * Build: see oracle/Makefile (must use -ffp-contract=off; FMAs are spelled fmaf()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_EINVAL 1

enum { ORC_KNN_APF_SQ = 0, ORC_KNN_P4P_CDIST = 1 };

static inline float sq3(float x, float y, float z) {
  float a = x * x, b = y * y, c = z * z;
  float s = a + b;
  return s + c;
}

/* sampler.py:4-30 furthest_point_sample / pix4point.py:8-53 farthest_point_sampling.
 * xyz: B clouds of N points, point p of cloud b at xyz[(b*N+p)*pt_stride + {0,1,2}].
 * start: the reference draws torch.randint(0,N,(B,)) (sampler.py:20); the caller passes it.
 * out: (B,G) int64.  No clamp of G here (the P4P flavour's min(n,N) is applied by callers). */
int orc_fps(const float* xyz, int64_t B, int64_t N, int64_t pt_stride, const int64_t* start,
            int64_t G, int64_t* out) {
  if (B < 0 || N <= 0 || G < 0 || pt_stride < 3) return ORC_EINVAL;
  int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < B; ++b) {
    const float* P = xyz + b * N * pt_stride;
    float* mind = (float*)malloc(sizeof(float) * (size_t)N);
    if (!mind) { err = 1; continue; }
    for (int64_t i = 0; i < N; ++i) mind[i] = 1e10f;
    int64_t far = start[b];
    if (far < 0 || far >= N) { err = 1; free(mind); continue; }
    for (int64_t g = 0; g < G; ++g) {
      out[b * G + g] = far;
      const float cx = P[far * pt_stride], cy = P[far * pt_stride + 1], cz = P[far * pt_stride + 2];
      float best = -1.0f;
      int64_t besti = 0;
      for (int64_t i = 0; i < N; ++i) {
        float dx = P[i * pt_stride] - cx, dy = P[i * pt_stride + 1] - cy, dz = P[i * pt_stride + 2] - cz;
        float d = sq3(dx, dy, dz);
        float m = mind[i];
        if (d < m) { m = d; mind[i] = m; }
        if (m > best) { best = m; besti = i; } /* strict >: first (lowest) index wins */
      }
      far = besti;
    }
    free(mind);
  }
  return err ? ORC_EINVAL : ORC_OK;
}

/* torch.sum(v, -1) over a CONTIGUOUS last dimension of D fp32 values on the CPU, restated (torch's cascade_sum,
 * aten/src/ATen/native/cpu/SumKernel.cpp; third-party, not under /root/reference).  Vector width 8 (the AVX2 kernel, also
 * what an AVX-512 host dispatches to).  D < 8 is the scalar row sum with four interleaved partial sums: p[j] = v[j] for
 * j < 4 when D >= 4, the elements from 4*(D/4) on are added to p[0], then p[0] += p[1], p[2], p[3].  D >= 8: lane k of the
 * vector accumulator = v[k] + v[8+k] + ... over the D/8 whole vectors, the tail elements are summed first (from 0), then
 * the 8 lanes are added in order.  D <= 4 and D = 8 come out as the plain left-to-right sum.  Checked bit for bit against
 * torch.sum for every D in 1..32 (tests/test_oracle_vs_reference.py::test_fps_nd).  q: D values; D <= 32. */
static float torch_cpu_row_sum(const float* q, int D) {
  if (D < 8) {
    float p[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    int done = 0;
    if (D >= 4) {
      for (int j = 0; j < 4; ++j) p[j] = p[j] + q[j];
      done = 4;
    }
    for (int i = done; i < D; ++i) p[0] = p[0] + q[i];
    for (int j = 1; j < 4; ++j) p[0] = p[0] + p[j];
    return p[0];
  }
  const int nvec = D / 8;
  float lane[8];
  for (int k = 0; k < 8; ++k) lane[k] = 0.0f + q[k];
  for (int v = 1; v < nvec; ++v)
    for (int k = 0; k < 8; ++k) lane[k] = lane[k] + q[8 * v + k];
  float f = 0.0f;
  for (int i = nvec * 8; i < D; ++i) f = f + q[i];
  for (int k = 0; k < 8; ++k) f = f + lane[k];
  return f;
}

/* rows x D values -> the row sums in torch's CPU order (test hook: compared with torch.sum bit for bit) */
int orc_torch_row_sum(const float* q, int64_t rows, int64_t D, float* out) {
  if (D < 1 || D > 32) return ORC_EINVAL;
  for (int64_t r = 0; r < rows; ++r) out[r] = torch_cpu_row_sum(q + r * D, (int)D);
  return ORC_OK;
}

/* pix4point.py:8-53 farthest_point_sampling on D-dimensional points (the distance sums over ALL D coordinates, line 44:
 * torch.sum((points - centroid) ** 2, -1)), 1 <= D <= 32; point p of cloud b at pts[(b*N+p)*pt_stride + 0..D-1].
 * The squared differences are summed in torch's CPU order (torch_cpu_row_sum).  The running distance starts at 1e10
 * (pix4point.py:27); update where dist < distance (47-48); torch.max keeps the first (lowest) index of the maximum (51).
 * D = 3 gives orc_fps's picks (the tests cross-check).  No clamp of G here (min(n_samples, N), line 23: callers). */
int orc_fps_nd(const float* pts, int64_t B, int64_t N, int64_t D, int64_t pt_stride, const int64_t* start,
               int64_t G, int64_t* out) {
  if (B < 0 || N <= 0 || G < 0 || D < 1 || D > 32 || pt_stride < D) return ORC_EINVAL;
  int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < B; ++b) {
    const float* P = pts + b * N * pt_stride;
    float* mind = (float*)malloc(sizeof(float) * (size_t)N);
    if (!mind) { err = 1; continue; }
    for (int64_t i = 0; i < N; ++i) mind[i] = 1e10f;
    int64_t far = start[b];
    if (far < 0 || far >= N) { err = 1; free(mind); continue; }
    for (int64_t g = 0; g < G; ++g) {
      out[b * G + g] = far;
      const float* c = P + far * pt_stride;
      float best = -1.0f;
      int64_t besti = 0;
      for (int64_t i = 0; i < N; ++i) {
        float q[32];
        for (int64_t a = 0; a < D; ++a) {
          float df = P[i * pt_stride + a] - c[a];
          q[a] = df * df;
        }
        float d = torch_cpu_row_sum(q, (int)D);
        float m = mind[i];
        if (d < m) { m = d; mind[i] = m; }
        if (m > best) { best = m; besti = i; } /* strict >: first (lowest) index wins */
      }
      far = besti;
    }
    free(mind);
  }
  return err ? ORC_EINVAL : ORC_OK;
}

static inline uint32_t f2ord(float f) {
  uint32_t b;
  memcpy(&b, &f, 4);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

static inline float knn_dist(int mode, float cx, float cy, float cz, float cn, float px, float py,
                             float pz) {
  float pn = sq3(px, py, pz);
  float d;
  if (mode == ORC_KNN_APF_SQ) {
    float dot = fmaf(cz, pz, fmaf(cy, py, cx * px));
    float t = -2.0f * dot;
    t = t + cn;
    d = t + pn;
  } else {
    float t = (-2.0f * cx) * px;
    t = fmaf(-2.0f * cy, py, t);
    t = fmaf(-2.0f * cz, pz, t);
    t = fmaf(cn, 1.0f, t);
    t = fmaf(1.0f, pn, t);
    t = t < 0.0f ? 0.0f : t;
    d = sqrtf(t);
  }
  return d + 0.0f; /* canonicalise -0.0 */
}

static void heap_sift_down(uint64_t* h, int64_t n, int64_t i) {
  for (;;) {
    int64_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < n && h[l] > h[m]) m = l;
    if (r < n && h[r] > h[m]) m = r;
    if (m == i) return;
    uint64_t t = h[i]; h[i] = h[m]; h[m] = t;
    i = m;
  }
}

static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

/* sampler.py:47-75 (_square_distance + knn_point, mode 0) and pix4point.py:79-89 (cdist+topk,
 * mode 1).  ctr: (B,G,3) contiguous.  idx_out (B,G,k) int64 ascending by (distance, index);
 * dist_out (B,G,k) float or NULL. */
int orc_knn(const float* xyz, int64_t pt_stride, const float* ctr, int64_t B, int64_t N, int64_t G,
            int64_t k, int mode, int64_t* idx_out, float* dist_out) {
  if (k <= 0 || k > N || pt_stride < 3 || (mode != 0 && mode != 1)) return ORC_EINVAL;
#pragma omp parallel for schedule(dynamic, 8)
  for (int64_t bg = 0; bg < B * G; ++bg) {
    int64_t b = bg / G;
    const float* P = xyz + b * N * pt_stride;
    const float cx = ctr[bg * 3], cy = ctr[bg * 3 + 1], cz = ctr[bg * 3 + 2];
    const float cn = sq3(cx, cy, cz);
    uint64_t* heap = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)k);
    int64_t hn = 0;
    for (int64_t i = 0; i < N; ++i) {
      float d = knn_dist(mode, cx, cy, cz, cn, P[i * pt_stride], P[i * pt_stride + 1],
                         P[i * pt_stride + 2]);
      uint64_t key = ((uint64_t)f2ord(d) << 32) | (uint64_t)(uint32_t)i;
      if (hn < k) {
        heap[hn++] = key;
        if (hn == k)
          for (int64_t j = k / 2 - 1; j >= 0; --j) heap_sift_down(heap, k, j);
      } else if (key < heap[0]) {
        heap[0] = key;
        heap_sift_down(heap, k, 0);
      }
    }
    qsort(heap, (size_t)k, sizeof(uint64_t), cmp_u64);
    for (int64_t j = 0; j < k; ++j) {
      int64_t i = (int64_t)(heap[j] & 0xffffffffu);
      idx_out[bg * k + j] = i;
      if (dist_out)
        dist_out[bg * k + j] = knn_dist(mode, cx, cy, cz, cn, P[i * pt_stride],
                                        P[i * pt_stride + 1], P[i * pt_stride + 2]);
    }
    free(heap);
  }
  return ORC_OK;
}

/* Full (B,G,N) distance matrix in either mode, for tie-window checks in tests. */
int orc_pair_dist(const float* xyz, int64_t pt_stride, const float* ctr, int64_t B, int64_t N,
                  int64_t G, int mode, float* out) {
  if (pt_stride < 3 || (mode != 0 && mode != 1)) return ORC_EINVAL;
#pragma omp parallel for schedule(static)
  for (int64_t bg = 0; bg < B * G; ++bg) {
    int64_t b = bg / G;
    const float* P = xyz + b * N * pt_stride;
    const float cx = ctr[bg * 3], cy = ctr[bg * 3 + 1], cz = ctr[bg * 3 + 2];
    const float cn = sq3(cx, cy, cz);
    for (int64_t i = 0; i < N; ++i)
      out[bg * N + i] = knn_dist(mode, cx, cy, cz, cn, P[i * pt_stride], P[i * pt_stride + 1],
                                 P[i * pt_stride + 2]);
  }
  return ORC_OK;
}

/* apf_utils.py:34-48 part1by2_vectorized, on int64 like the reference. */
static inline int64_t part1by2(int64_t n) {
  n = n & 0x000003ff;
  n = (n ^ (n << 16)) & 0xff0000ff;
  n = (n ^ (n << 8)) & 0x0300f00f;
  n = (n ^ (n << 4)) & 0x030c30c3;
  n = (n ^ (n << 2)) & 0x09249249;
  return n;
}

/* apf_utils.py:66-104 points_to_morton (resolution=1024): per-cloud min/max over the G centres
 * (89-90), (p-min)/((max-min)+1e-8f) (91), *1023 then truncation to int64 (92), 10-bit
 * interleave z<<2 + y<<1 + x (62-64).  codes: (B,G) int64.  perm: (B,G) int64 = STABLE
 * ascending argsort of the codes (canonical; torch.argsort at 103 is not stable). */
int orc_morton(const float* ctr, int64_t B, int64_t G, int64_t* codes, int64_t* perm) {
  if (G <= 0) return ORC_EINVAL;
  for (int64_t b = 0; b < B; ++b) {
    const float* C = ctr + b * G * 3;
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) { mn[a] = C[a]; mx[a] = C[a]; }
    for (int64_t g = 1; g < G; ++g)
      for (int a = 0; a < 3; ++a) {
        float v = C[g * 3 + a];
        if (v < mn[a]) mn[a] = v;
        if (v > mx[a]) mx[a] = v;
      }
    for (int64_t g = 0; g < G; ++g) {
      int64_t q[3];
      for (int a = 0; a < 3; ++a) {
        float num = C[g * 3 + a] - mn[a];
        float den = mx[a] - mn[a];
        den = den + 1e-8f;
        float nrm = num / den;
        float sc = nrm * 1023.0f;
        q[a] = (int64_t)sc;
      }
      codes[b * G + g] = (part1by2(q[2]) << 2) + (part1by2(q[1]) << 1) + part1by2(q[0]);
    }
    if (perm) {
      /* stable insertion-free: sort (code<<32 | g) keys; codes < 2^30, g < 2^31 */
      uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)G);
      for (int64_t g = 0; g < G; ++g) keys[g] = ((uint64_t)codes[b * G + g] << 32) | (uint64_t)g;
      qsort(keys, (size_t)G, sizeof(uint64_t), cmp_u64);
      for (int64_t g = 0; g < G; ++g) perm[b * G + g] = (int64_t)(keys[g] & 0xffffffffu);
      free(keys);
    }
  }
  return ORC_OK;
}

/* apf.py:52-112 Group.forward, given the index results: x (B,N,C) contiguous, fps_idx (B,G),
 * knn_idx (B,G,k), perm (B,G) (Morton order; output group j of cloud b = group perm[b,j]).
 * neigh: (B,G,k,2C) = [x[nbr]-x[centre] || x[centre]]; center: (B,G,3) = x[centre,:3]. */
int orc_group_apf(const float* x, int64_t B, int64_t N, int64_t C, const int64_t* fps_idx,
                  const int64_t* knn_idx, const int64_t* perm, int64_t G, int64_t k, float* neigh,
                  float* center) {
  if (C < 3) return ORC_EINVAL;
  for (int64_t b = 0; b < B; ++b)
    for (int64_t j = 0; j < G; ++j) {
      int64_t g = perm ? perm[b * G + j] : j;
      const float* cr = x + (b * N + fps_idx[b * G + g]) * C;
      for (int a = 0; a < 3; ++a) center[(b * G + j) * 3 + a] = cr[a];
      for (int64_t n = 0; n < k; ++n) {
        const float* pr = x + (b * N + knn_idx[(b * G + g) * k + n]) * C;
        float* o = neigh + ((b * G + j) * k + n) * 2 * C;
        for (int64_t c = 0; c < C; ++c) {
          o[c] = pr[c] - cr[c];
          o[C + c] = cr[c];
        }
      }
    }
  return ORC_OK;
}
