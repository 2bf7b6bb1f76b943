// lm_icp.cu -- poseRefine (point-to-plane ICP pose refinement) for sm_100a + its C-ABI.
//
// Reference being replaced: poseRefine::process, linemodLevelup/linemodLevelup.cpp:27-155 of
// meiqua/6DPose @ 619be57 ("LL.cpp"), and the Open3D calls it makes (Open3D is external and unpinned
// in the reference -- oracle/icp_oracle.py documents the restated semantics; PARITY UNPINNED).
//   host (this file)      <- mask dilation + bounding box, cloud construction, voxel down-sampling
//                            (LL.cpp:34-109; PointCloud::VoxelDownSample)
//   k_icp_normals         <- PointCloud::EstimateNormals, KNN = 30 (LL.cpp:127)
//   k_icp_point_to_plane  <- RegistrationICP + TransformationEstimationPointToPlane (LL.cpp:128-130):
//                            nearest-neighbour correspondence search fused with the 6x6 J^T J / J^T r
//                            accumulation, the 6x6 solve and the convergence test; every iteration of
//                            one hypothesis runs inside one CTA, one launch for a batch of hypotheses
// All arithmetic is double, as in Open3D/Eigen.  The reference's quirks are kept: the ICP target is the
// down-sampled MODEL cloud (LL.cpp:109), only the z component of the initial translation is converted
// to metres (LL.cpp:37), `residual` carries the fitness (LL.cpp:148).

#include "linemod_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <thread>
#include <vector>

int lm_fail(int code, const char* fmt, ...);  // linemod_b200.cu

#define CUI(call)                                                                                      \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return lm_fail(LM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define ICP_KNN 30
#define ICP_TILE 256
#define ICP_THREADS 512
#define ICP_NACC 30  // 21 (upper J^T J) + 6 (J^T r) + squared error + correspondences + 1 spare

struct IcpJob {  // one hypothesis
  int32_t src_first, src_count;  // source cloud: range in the concatenated point array
  int32_t tgt_first, tgt_count;  // target cloud (+ normals); the reference uses the source cloud again
  double init[16];               // row-major 4x4 initial guess
};

struct IcpOut {
  double T[16];  // row-major final transformation_
  double fitness, rmse;
  int32_t iterations, pad;
};

// --------------------------------------------------------------------------------------------
// normals: covariance of the 30 nearest neighbours, eigenvector of the smallest eigenvalue
// --------------------------------------------------------------------------------------------
__device__ void smallest_eigenvector_3x3(double a00, double a01, double a02, double a11, double a12, double a22, double* n) {
  // cyclic Jacobi on the symmetric 3x3 matrix, eigenvectors accumulated in v
  double a[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 24; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int m = 0;
  if (a[1][1] < a[m][m]) m = 1;
  if (a[2][2] < a[m][m]) m = 2;
  const double nx = v[0][m], ny = v[1][m], nz = v[2][m];
  const double len = sqrt(nx * nx + ny * ny + nz * nz);
  if (len > 0) { n[0] = nx / len; n[1] = ny / len; n[2] = nz / len; }
  else { n[0] = 0; n[1] = 0; n[2] = 1; }
}

// The clouds come out of voxel_down_sample sorted by voxel (x major), so index distance is spatial distance along
// x: tiles are visited outwards from the CTA's own points, and a side is abandoned once the x gap alone puts every
// remaining candidate beyond the current 30th neighbour of every thread (exact: a point of voxel column a has
// x in [lo_a, lo_a + ICP_VOXEL)).  The kept set is the 30 smallest (distance, index) pairs in that order --
// the same list a scan in index order with "first seen wins" produces -- so the normals do not depend on the
// visiting order.
#define ICP_VOXEL 0.0025

__device__ __forceinline__ void knn_scan_tile(const double (*s_t)[3], int t0, int m, double px, double py, double pz,
                                              double* bd, int* bi, int& have) {
  for (int k = 0; k < m; ++k) {
    const double dx = px - s_t[k][0], dy = py - s_t[k][1], dz = pz - s_t[k][2];
    const double d2 = dx * dx + dy * dy + dz * dz;
    const int idx = t0 + k;
    if (have < ICP_KNN || d2 < bd[have - 1] || (d2 == bd[have - 1] && idx < bi[have - 1])) {
      int pos = have < ICP_KNN ? have++ : ICP_KNN - 1;
      while (pos > 0 && (bd[pos - 1] > d2 || (bd[pos - 1] == d2 && bi[pos - 1] > idx))) {
        bd[pos] = bd[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
      }
      bd[pos] = d2;
      bi[pos] = idx;
    }
  }
}

__global__ void __launch_bounds__(128) k_icp_normals(const IcpJob* __restrict__ jobs, const double* __restrict__ pts,
                                                     double* __restrict__ normals) {
  __shared__ double s_t[ICP_TILE][3];
  const IcpJob job = jobs[blockIdx.y];
  const int n = job.tgt_count;
  const double* __restrict__ P = pts + (size_t)job.tgt_first * 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  double px = 0, py = 0, pz = 0;
  if (live) { px = P[3 * i]; py = P[3 * i + 1]; pz = P[3 * i + 2]; }
  double bd[ICP_KNN];
  int bi[ICP_KNN];
  int have = 0;
  const int ntiles = (n + ICP_TILE - 1) / ICP_TILE;
  if (ntiles > 0) {
    const int center = min((int)((blockIdx.x * blockDim.x + blockDim.x / 2) / ICP_TILE), ntiles - 1);
    int lo = center - 1, hi = center + 1;
    bool more_lo = lo >= 0, more_hi = hi < ntiles;
    int tile = center;
    for (;;) {
      const int t0 = tile * ICP_TILE;
      __syncthreads();
      for (int k = threadIdx.x; k < ICP_TILE * 3; k += blockDim.x) {
        const int j = t0 + k / 3;
        (&s_t[0][0])[k] = j < n ? P[(size_t)3 * t0 + k] : 0.0;
      }
      __syncthreads();
      if (live) knn_scan_tile(s_t, t0, min(ICP_TILE, n - t0), px, py, pz, bd, bi, have);
      // next tile: alternate sides while they are worth visiting (block-uniform decisions)
      bool picked = false;
      for (int side = 0; side < 2 && !picked; ++side) {
        const bool up = ((tile >= center) == (side == 0)) ? false : true;  // prefer the side not just visited
        if (up && more_hi) {
          const double gap = P[(size_t)3 * hi * ICP_TILE] - ICP_VOXEL - px;  // every x beyond is > this
          const bool useless = !live || (have == ICP_KNN && gap > 0.0 && gap * gap > bd[ICP_KNN - 1]);
          if (__syncthreads_and(useless)) more_hi = false;
          else { tile = hi++; more_hi = hi < ntiles; picked = true; }
        } else if (!up && more_lo) {
          const int last = min((lo + 1) * ICP_TILE, n) - 1;
          const double gap = px - (P[(size_t)3 * last] + ICP_VOXEL);  // every x before is < that
          const bool useless = !live || (have == ICP_KNN && gap > 0.0 && gap * gap > bd[ICP_KNN - 1]);
          if (__syncthreads_and(useless)) more_lo = false;
          else { tile = lo--; more_lo = lo >= 0; picked = true; }
        }
      }
      if (!picked) break;
    }
  }
  if (!live) return;
  double* out = normals + ((size_t)job.tgt_first + i) * 3;
  if (have < 3) { out[0] = 0; out[1] = 0; out[2] = 1; return; }
  // cumulants exactly as Open3D's ComputeNormal: E[x], E[x x^T], covariance = E[x x^T] - E[x] E[x]^T
  double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < have; ++k) {
    const double x = P[3 * bi[k]], y = P[3 * bi[k] + 1], z = P[3 * bi[k] + 2];
    c[0] += x; c[1] += y; c[2] += z;
    c[3] += x * x; c[4] += x * y; c[5] += x * z; c[6] += y * y; c[7] += y * z; c[8] += z * z;
  }
  for (int k = 0; k < 9; ++k) c[k] /= (double)have;
  smallest_eigenvector_3x3(c[3] - c[0] * c[0], c[4] - c[0] * c[1], c[5] - c[0] * c[2], c[6] - c[1] * c[1], c[7] - c[1] * c[2],
                           c[8] - c[2] * c[2], out);
}

// --------------------------------------------------------------------------------------------
// ICP: all iterations of one hypothesis inside one CTA
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// Solve the symmetric 6x6 system A x = b (Gaussian elimination with partial pivoting; Open3D uses
// Eigen's LDLT, same solution to rounding).  A singular system gives x = 0 (identity update).
__device__ void solve6(double A[6][6], double b[6], double x[6]) {
  bool singular = false;
  for (int k = 0; k < 6; ++k) {
    int m = k;
    for (int i = k + 1; i < 6; ++i)
      if (fabs(A[i][k]) > fabs(A[m][k])) m = i;
    if (!(fabs(A[m][k]) > 0.0)) { singular = true; break; }
    if (m != k) {
      for (int j = 0; j < 6; ++j) { const double t = A[k][j]; A[k][j] = A[m][j]; A[m][j] = t; }
      const double t = b[k]; b[k] = b[m]; b[m] = t;
    }
    for (int i = k + 1; i < 6; ++i) {
      const double f = A[i][k] / A[k][k];
      for (int j = k; j < 6; ++j) A[i][j] -= f * A[k][j];
      b[i] -= f * b[k];
    }
  }
  if (!singular)
    for (int i = 5; i >= 0; --i) {
      double s = b[i];
      for (int j = i + 1; j < 6; ++j) s -= A[i][j] * x[j];
      x[i] = s / A[i][i];
    }
  bool ok = !singular;
  for (int i = 0; i < 6 && ok; ++i) ok = isfinite(x[i]);
  if (!ok)
    for (int i = 0; i < 6; ++i) x[i] = 0.0;
}

__global__ void __launch_bounds__(ICP_THREADS) k_icp_point_to_plane(const IcpJob* __restrict__ jobs, const double* __restrict__ pts,
                                                                  const double* __restrict__ normals, IcpOut* __restrict__ outs,
                                                                  double max_d2, int max_iter, double rel_fitness, double rel_rmse) {
  __shared__ double s_t[ICP_TILE][3];
  __shared__ double s_red[ICP_THREADS / 32][ICP_NACC];
  __shared__ double s_T[16];
  __shared__ int s_stop;
  const IcpJob job = jobs[blockIdx.x];
  const int n = job.src_count, nt = job.tgt_count;
  const double* __restrict__ P = pts + (size_t)job.src_first * 3;  // source cloud
  const double* __restrict__ Q = pts + (size_t)job.tgt_first * 3;  // target cloud (== source in the reference, LL.cpp:108-109)
  const double* __restrict__ Nm = normals + (size_t)job.tgt_first * 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 16) s_T[threadIdx.x] = job.init[threadIdx.x];
  if (threadIdx.x == 0) s_stop = 0;
  __syncthreads();

  double fit_prev = 0, rmse_prev = 0, fitness = 0, rmse = 0;
  int it = 0;
  for (int pass = 0;; ++pass) {
    const double r00 = s_T[0], r01 = s_T[1], r02 = s_T[2], tx = s_T[3];
    const double r10 = s_T[4], r11 = s_T[5], r12 = s_T[6], ty = s_T[7];
    const double r20 = s_T[8], r21 = s_T[9], r22 = s_T[10], tz = s_T[11];
    double acc[ICP_NACC];
#pragma unroll
    for (int k = 0; k < ICP_NACC; ++k) acc[k] = 0.0;

    for (int c0 = 0; c0 < n; c0 += ICP_THREADS) {
      const int i = c0 + threadIdx.x;
      const bool live = i < n;
      double px = 0, py = 0, pz = 0;
      if (live) {
        const double x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
        px = r00 * x + r01 * y + r02 * z + tx;
        py = r10 * x + r11 * y + r12 * z + ty;
        pz = r20 * x + r21 * y + r22 * z + tz;
      }
      double best = max_d2;  // radius search, strict <
      int bj = -1;
      for (int t0 = 0; t0 < nt; t0 += ICP_TILE) {
        // The target cloud is sorted by voxel, x major (voxel_down_sample), so a tile covers the x range
        // [x(first) - voxel, x(last) + voxel]: a tile whose range is farther from EVERY source point of this pass than its
        // best squared distance so far cannot change any correspondence (strict <) and is skipped -- exact, and most
        // tiles go: the search radius is 1 cm, the cloud is ~10 cm wide
        {
          const int last = min(t0 + ICP_TILE, nt) - 1;
          const double tmin = Q[(size_t)3 * t0] - ICP_VOXEL, tmax = Q[(size_t)3 * last] + ICP_VOXEL;
          const double gap = !live ? 1e300 : (px < tmin ? tmin - px : (px > tmax ? px - tmax : 0.0));
          if (__syncthreads_and(!live || gap * gap >= best)) continue;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < ICP_TILE * 3; k += blockDim.x) (&s_t[0][0])[k] = (t0 + k / 3) < nt ? Q[(size_t)3 * t0 + k] : 0.0;
        __syncthreads();
        if (live) {
          const int m = min(ICP_TILE, nt - t0);
#pragma unroll 4
          for (int k = 0; k < m; ++k) {
            const double dx = px - s_t[k][0], dy = py - s_t[k][1], dz = pz - s_t[k][2];
            const double d2 = dx * dx + dy * dy + dz * dz;
            if (d2 < best) { best = d2; bj = t0 + k; }
          }
        }
      }
      if (bj >= 0) {
        const double vx = Q[3 * bj], vy = Q[3 * bj + 1], vz = Q[3 * bj + 2];
        const double nx = Nm[3 * bj], ny = Nm[3 * bj + 1], nz = Nm[3 * bj + 2];
        const double r = (px - vx) * nx + (py - vy) * ny + (pz - vz) * nz;
        const double J[6] = {py * nz - pz * ny, pz * nx - px * nz, px * ny - py * nx, nx, ny, nz};
        int q = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = a; b < 6; ++b) acc[q++] += J[a] * J[b];
#pragma unroll
        for (int a = 0; a < 6; ++a) acc[21 + a] += J[a] * r;
        acc[27] += best;
        acc[28] += 1.0;
      }
    }
    // block reduction of the 29 sums
#pragma unroll
    for (int k = 0; k < ICP_NACC; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0)
#pragma unroll
      for (int k = 0; k < ICP_NACC; ++k) s_red[warp][k] = acc[k];
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot[ICP_NACC];
      for (int k = 0; k < ICP_NACC; ++k) {
        double s = 0;
        for (int w = 0; w < ICP_THREADS / 32; ++w) s += s_red[w][k];
        tot[k] = s;
      }
      const double ncorr = tot[28];
      fitness = n > 0 ? ncorr / (double)n : 0.0;
      rmse = ncorr > 0 ? sqrt(tot[27] / ncorr) : 0.0;
      bool stop = false;
      if (pass > 0 && fabs(fit_prev - fitness) < rel_fitness && fabs(rmse_prev - rmse) < rel_rmse) stop = true;
      if (pass >= max_iter) stop = true;
      if (!stop) {
        double A[6][6], b[6], x[6] = {0, 0, 0, 0, 0, 0};
        int q = 0;
        for (int a = 0; a < 6; ++a)
          for (int c = a; c < 6; ++c) { A[a][c] = tot[q]; A[c][a] = tot[q]; ++q; }
        for (int a = 0; a < 6; ++a) b[a] = -tot[21 + a];
        solve6(A, b, x);
        // TransformVector6dToMatrix4d: Rz(x2) Ry(x1) Rx(x0), translation x3..5; T = update * T
        const double cx = cos(x[0]), sx = sin(x[0]), cy = cos(x[1]), sy = sin(x[1]), cz = cos(x[2]), sz = sin(x[2]);
        const double U[12] = {cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx, x[3],
                              sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx, x[4],
                              -sy,     cy * sx,                cy * cx,                x[5]};
        double Tn[12];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 4; ++c)
            Tn[4 * r + c] = U[4 * r] * s_T[c] + U[4 * r + 1] * s_T[4 + c] + U[4 * r + 2] * s_T[8 + c] + (c == 3 ? U[4 * r + 3] : 0.0);
        for (int k = 0; k < 12; ++k) s_T[k] = Tn[k];
        fit_prev = fitness;
        rmse_prev = rmse;
        it = pass + 1;
      }
      s_stop = stop ? 1 : 0;
    }
    __syncthreads();
    if (s_stop) break;
  }
  if (threadIdx.x == 0) {
    IcpOut o;
    for (int k = 0; k < 16; ++k) o.T[k] = s_T[k];
    o.fitness = fitness; o.rmse = rmse; o.iterations = it; o.pad = 0;
    outs[blockIdx.x] = o;
  }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
struct lm_icp {
  int device = 0;
  cudaStream_t stream = nullptr;
  double* d_pts = nullptr; double* d_nrm = nullptr; size_t cap_pts = 0;
  IcpJob* d_jobs = nullptr; IcpOut* d_outs = nullptr; size_t cap_jobs = 0;
  IcpOut* h_outs = nullptr; size_t cap_houts = 0;
  int64_t launches = 0;
  bool use_scene_cloud = false;  // non-reference option: register against the scene cloud
  int last_points = 0, last_iterations = 0;
  double last_rmse = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_kernel_us = 0;
};

extern "C" int lm_icp_create(int device, lm_icp** out) {
  if (!out) return lm_fail(LM_E_INVALID, "out is null");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return lm_fail(LM_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return lm_fail(LM_E_INVALID, "device %d out of range (have %d)", device, ndev);
  CUI(cudaSetDevice(device));
  lm_icp* h = new lm_icp();
  h->device = device;
  CUI(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CUI(cudaEventCreate(&h->ev0));
  CUI(cudaEventCreate(&h->ev1));
  *out = h;
  return LM_OK;
}

extern "C" void lm_icp_destroy(lm_icp* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaFree(h->d_pts); cudaFree(h->d_nrm); cudaFree(h->d_jobs); cudaFree(h->d_outs);
  cudaFreeHost(h->h_outs);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
  cudaStreamDestroy(h->stream);
  delete h;
}

namespace {

struct Cloud {
  std::vector<double> pts;    // down-sampled model cloud, xyz
  std::vector<double> scene;  // down-sampled scene cloud (only built when asked for)
  double init[16];
  float base[16];  // init_base, row-major
  bool early = false;
};

// LL.cpp:34-109 for one hypothesis.
void voxel_down_sample(const std::vector<double>& mp, std::vector<double>& out);

void build_cloud(const uint16_t* scene, int srows, int scols, const uint16_t* model, int mrows, int mcols, const float* sK,
                 const float* mK, const float* R, const float* t, int detectX, int detectY, bool want_scene, Cloud& c) {
  // init_base = [R | t] with only t.z converted to metres (LL.cpp:34-38)
  for (int i = 0; i < 16; ++i) c.base[i] = 0.f;
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) c.base[4 * r + k] = R[3 * r + k];
    c.base[4 * r + 3] = t[r];
  }
  c.base[4 * 2 + 3] /= 1000.0f;
  c.base[15] = 1.f;
  // dilate(modelDepth > 0, 9x9) + boundingRect (LL.cpp:43-50): the box of the non-zero pixels grown by 4
  const int half = 4;
  int x0 = mcols, x1 = -1, y0 = mrows, y1 = -1;
  for (int r = 0; r < mrows; ++r)
    for (int q = 0; q < mcols; ++q)
      if (model[(size_t)r * mcols + q] > 0) { x0 = std::min(x0, q); x1 = std::max(x1, q); y0 = std::min(y0, r); y1 = std::max(y1, r); }
  int bx = 0, by = 0, bw = 0, bh = 0;
  if (x1 >= 0) {
    bx = std::max(x0 - half, 0); by = std::max(y0 - half, 0);
    bw = std::min(x1 + half, mcols - 1) - bx + 1;
    bh = std::min(y1 + half, mrows - 1) - by + 1;
  }
  c.early = (detectX + bw >= scols) || (detectY + bh >= srows);  // LL.cpp:52-55
  if (c.early) return;
  // clouds (LL.cpp:57-99)
  std::vector<double> mp, sp;
  double cm[3] = {0, 0, 0}, cs[3] = {0, 0, 0};
  long ncs = 0;
  const double anchor = model[(size_t)(mrows / 2) * mcols + mcols / 2] / 1000.0;
  auto in_mask = [&](int mr, int mc) {  // dilated mask
    for (int dy = -half; dy <= half; ++dy) {
      const int yy = mr + dy;
      if (yy < 0 || yy >= mrows) continue;
      for (int dx = -half; dx <= half; ++dx) {
        const int xx = mc + dx;
        if (xx >= 0 && xx < mcols && model[(size_t)yy * mcols + xx] > 0) return true;
      }
    }
    return false;
  };
  for (int r = 0; r < bh; ++r)
    for (int q = 0; q < bw; ++q) {
      const int mr = r + by, mc = q + bx;
      int sr = r + detectY - half; if (sr < 0) sr = 0;
      int sc = q + detectX - half; if (sc < 0) sc = 0;
      const uint16_t md = model[(size_t)mr * mcols + mc];
      if (md == 0 && !in_mask(mr, mc)) continue;
      if (md > 0) {
        const double z = md / 1000.0;
        const double x = (mc - mK[2]) / mK[0] * z;  // int - float -> float, / float, * double
        const double y = (mr - mK[5]) / mK[4] * z;
        mp.push_back(x); mp.push_back(y); mp.push_back(z);
        cm[0] += x; cm[1] += y; cm[2] += z;
      }
      const uint16_t sd = scene[(size_t)sr * scols + sc];
      if (sd > 0) {
        const double z = sd / 1000.0;
        const double x = (sc - sK[2]) / sK[0] * z;
        const double y = (sr - sK[5]) / sK[4] * z;
        if (want_scene) { sp.push_back(x); sp.push_back(y); sp.push_back(z); }
        if (fabs(z - anchor) < 0.4 && md > 0) { cs[0] += x; cs[1] += y; cs[2] += z; ++ncs; }
      }
    }
  const double nm = (double)(mp.size() / 3);
  for (int k = 0; k < 3; ++k) { cm[k] /= nm; cs[k] /= (double)ncs; }  // 0/0 -> NaN as in the reference
  for (int i = 0; i < 16; ++i) c.init[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int k = 0; k < 3; ++k) c.init[4 * k + 3] = cs[k] - cm[k];
  voxel_down_sample(mp, c.pts);
  if (want_scene) voxel_down_sample(sp, c.scene);
}

// PointCloud::VoxelDownSample(0.0025) (LL.cpp:106-109): mean of the points of every occupied voxel;
// Open3D emits them in hash-map order, here in sorted voxel order (affects summation order only).
void voxel_down_sample(const std::vector<double>& mp, std::vector<double>& out) {
  const double voxel = 0.0025;
  const size_t n = mp.size() / 3;
  out.clear();
  if (n == 0) return;
  double mn[3] = {mp[0], mp[1], mp[2]};
  for (size_t i = 1; i < n; ++i)
    for (int k = 0; k < 3; ++k) mn[k] = std::min(mn[k], mp[3 * i + k]);
  struct Key { int64_t a, b, c; size_t i; };
  std::vector<Key> keys(n);
  for (size_t i = 0; i < n; ++i) {
    keys[i].a = (int64_t)floor((mp[3 * i] - (mn[0] - voxel * 0.5)) / voxel);
    keys[i].b = (int64_t)floor((mp[3 * i + 1] - (mn[1] - voxel * 0.5)) / voxel);
    keys[i].c = (int64_t)floor((mp[3 * i + 2] - (mn[2] - voxel * 0.5)) / voxel);
    keys[i].i = i;
  }
  std::stable_sort(keys.begin(), keys.end(), [](const Key& p, const Key& q) {
    if (p.a != q.a) return p.a < q.a;
    if (p.b != q.b) return p.b < q.b;
    return p.c < q.c;
  });
  for (size_t s = 0; s < n;) {
    size_t e = s;
    double acc[3] = {0, 0, 0};
    while (e < n && keys[e].a == keys[s].a && keys[e].b == keys[s].b && keys[e].c == keys[s].c) {
      for (int k = 0; k < 3; ++k) acc[k] += mp[3 * keys[e].i + k];
      ++e;
    }
    for (int k = 0; k < 3; ++k) out.push_back(acc[k] / (double)(e - s));
    s = e;
  }
}

}  // namespace

// Batch of hypotheses against one scene depth image.  For hypothesis h: model_depths[h] (mrows x mcols
// u16), modelK + 9h, R + 9h, t + 3h, detect_xy[2h..2h+1].  Outputs R_out + 9h (row-major f64),
// t_out + 3h (mm), residual[h] (fitness; -1 on the early return of LL.cpp:52-55, outputs untouched).
extern "C" int lm_icp_process_batch(lm_icp* h, int n_hyp, const uint16_t* scene_depth, int srows, int scols,
                                    const uint16_t* const* model_depths, int mrows, int mcols, const float* sceneK,
                                    const float* modelK, const float* R, const float* t, const int32_t* detect_xy,
                                    int max_iterations, double* R_out, double* t_out, float* residual) {
  if (!h || n_hyp < 0 || !scene_depth || !model_depths || !sceneK || !modelK || !R || !t || !detect_xy || !R_out || !t_out || !residual)
    return lm_fail(LM_E_INVALID, "null argument");
  if (srows <= 0 || scols <= 0 || mrows <= 0 || mcols <= 0) return lm_fail(LM_E_INVALID, "empty image");
  if (max_iterations < 0) return lm_fail(LM_E_INVALID, "max_iterations < 0");
  CUI(cudaSetDevice(h->device));
  std::vector<Cloud> clouds((size_t)n_hyp);
  std::vector<IcpJob> jobs;
  std::vector<int> job_of((size_t)n_hyp, -1);
  std::vector<double> all;
  int max_count = 0;
  for (int i = 0; i < n_hyp; ++i)
    if (!model_depths[i]) return lm_fail(LM_E_INVALID, "model_depths[%d] is null", i);
  {
    // the hypotheses' clouds are independent: hypothesis 0 on this thread, the others on their own (a batch is the three
    // NMS survivors of the drivers, linemod_and_levelup_test.py:348-367)
    auto build = [&](int i) {
      build_cloud(scene_depth, srows, scols, model_depths[i], mrows, mcols, sceneK, modelK + 9 * i, R + 9 * i, t + 3 * i,
                  detect_xy[2 * i], detect_xy[2 * i + 1], h->use_scene_cloud, clouds[i]);
    };
    std::vector<std::thread> workers;
    for (int i = 1; i < n_hyp && i < 16; ++i) workers.emplace_back(build, i);
    if (n_hyp > 0) build(0);
    for (int i = 16; i < n_hyp; ++i) build(i);
    for (std::thread& w : workers) w.join();
  }
  for (int i = 0; i < n_hyp; ++i) {
    if (clouds[i].early) continue;
    IcpJob j;
    j.src_first = (int32_t)(all.size() / 3);
    j.src_count = (int32_t)(clouds[i].pts.size() / 3);
    all.insert(all.end(), clouds[i].pts.begin(), clouds[i].pts.end());
    j.tgt_first = j.src_first;  // scene_pcd_down = model_pcd->VoxelDownSample (LL.cpp:109)
    j.tgt_count = j.src_count;
    if (h->use_scene_cloud) {
      j.tgt_first = (int32_t)(all.size() / 3);
      j.tgt_count = (int32_t)(clouds[i].scene.size() / 3);
      all.insert(all.end(), clouds[i].scene.begin(), clouds[i].scene.end());
    }
    memcpy(j.init, clouds[i].init, sizeof(j.init));
    job_of[i] = (int)jobs.size();
    jobs.push_back(j);
    max_count = std::max(max_count, j.tgt_count);
  }
  const size_t npts = all.size() / 3;
  if (!jobs.empty()) {
    if (npts > h->cap_pts) {
      cudaFree(h->d_pts); cudaFree(h->d_nrm);
      h->cap_pts = std::max<size_t>(npts * 2, 4096);
      CUI(cudaMalloc(&h->d_pts, h->cap_pts * 3 * sizeof(double)));
      CUI(cudaMalloc(&h->d_nrm, h->cap_pts * 3 * sizeof(double)));
    }
    if (jobs.size() > h->cap_jobs) {
      cudaFree(h->d_jobs); cudaFree(h->d_outs);
      h->cap_jobs = std::max<size_t>(jobs.size() * 2, 16);
      CUI(cudaMalloc(&h->d_jobs, h->cap_jobs * sizeof(IcpJob)));
      CUI(cudaMalloc(&h->d_outs, h->cap_jobs * sizeof(IcpOut)));
    }
    if (jobs.size() > h->cap_houts) {
      cudaFreeHost(h->h_outs);
      h->cap_houts = std::max<size_t>(jobs.size() * 2, 16);
      CUI(cudaMallocHost(&h->h_outs, h->cap_houts * sizeof(IcpOut)));
    }
    cudaStream_t st = h->stream;
    if (npts) CUI(cudaMemcpyAsync(h->d_pts, all.data(), npts * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    CUI(cudaMemcpyAsync(h->d_jobs, jobs.data(), jobs.size() * sizeof(IcpJob), cudaMemcpyHostToDevice, st));
    CUI(cudaEventRecord(h->ev0, st));
    if (max_count > 0) {
      dim3 g((unsigned)((max_count + 127) / 128), (unsigned)jobs.size());
      k_icp_normals<<<g, 128, 0, st>>>(h->d_jobs, h->d_pts, h->d_nrm);
      ++h->launches;
    }
    k_icp_point_to_plane<<<(unsigned)jobs.size(), ICP_THREADS, 0, st>>>(h->d_jobs, h->d_pts, h->d_nrm, h->d_outs, 0.01 * 0.01,
                                                                      max_iterations, 1e-6, 1e-6);
    ++h->launches;
    CUI(cudaEventRecord(h->ev1, st));
    CUI(cudaMemcpyAsync(h->h_outs, h->d_outs, jobs.size() * sizeof(IcpOut), cudaMemcpyDeviceToHost, st));
    CUI(cudaStreamSynchronize(st));
    CUI(cudaGetLastError());
    float ms = 0;
    CUI(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_kernel_us = ms * 1000.f;
  }
  for (int i = 0; i < n_hyp; ++i) {
    if (job_of[i] < 0) { residual[i] = -1.f; continue; }
    const IcpOut& o = h->h_outs[job_of[i]];
    // result = transformation_ * init_base (LL.cpp:146); R_refined, t_refined * 1000 (LL.cpp:153-154)
    double res[16];
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        double s = 0;
        for (int k = 0; k < 4; ++k) s += o.T[4 * r + k] * (double)clouds[i].base[4 * k + c];
        res[4 * r + c] = s;
      }
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) R_out[9 * i + 3 * r + c] = res[4 * r + c];
      t_out[3 * i + r] = res[4 * r + 3] * 1000.0;
    }
    residual[i] = (float)o.fitness;
    h->last_points = jobs[job_of[i]].src_count;
    h->last_iterations = o.iterations;
    h->last_rmse = o.rmse;
  }
  return LM_OK;
}

extern "C" int lm_icp_process(lm_icp* h, const uint16_t* scene_depth, int srows, int scols, const uint16_t* model_depth, int mrows,
                              int mcols, const float* sceneK, const float* modelK, const float* R, const float* t, int detectX,
                              int detectY, int max_iterations, double* R_out, double* t_out, float* residual) {
  const uint16_t* models[1] = {model_depth};
  const int32_t xy[2] = {detectX, detectY};
  return lm_icp_process_batch(h, 1, scene_depth, srows, scols, models, mrows, mcols, sceneK, modelK, R, t, xy, max_iterations, R_out,
                              t_out, residual);
}

// [0] points in the down-sampled cloud, [1] ICP iterations, [2] inlier rmse (m), [3] device time of the
// normals + ICP kernels of the last call in microseconds -- all of the LAST hypothesis processed.
extern "C" int lm_icp_last_stats(lm_icp* h, double* out4) {
  if (!h || !out4) return lm_fail(LM_E_INVALID, "null argument");
  out4[0] = h->last_points; out4[1] = h->last_iterations; out4[2] = h->last_rmse; out4[3] = h->last_kernel_us;
  return LM_OK;
}

extern "C" int64_t lm_icp_launch_count(lm_icp* h) { return h ? h->launches : 0; }

extern "C" int lm_icp_set_use_scene_cloud(lm_icp* h, int on) {
  if (!h) return lm_fail(LM_E_INVALID, "null handle");
  h->use_scene_cloud = on != 0;
  return LM_OK;
}
