// Post-match stage on the device.
//
// The reference's callers never use the raw match list: linemod_and_levelup_test.py:325-345 builds boxes
// (x, y, x + width, y + height, similarity) and runs nms(dets, 0.5) (:34-61), then refines the first three
// survivors with poseRefine (:348-367); linemod_ros/detect.py:41-81, 94-134 does the same.  Here the greedy
// NMS runs on the device right behind k_refine, on the kept records where they lie, and only the top_k
// survivors cross PCIe.
//
// Greedy NMS without a sort: round r picks the best live record (block-wide argmax), emits it, and kills every
// live record whose IoU with it exceeds the threshold -- exactly the reference's loop
// ("order = order[where(ovr <= thresh)]"), with the IoU in float64 and the "+1" pixel convention of :40, :51-54.
// Order among equal similarities, which the reference leaves to an unstable std::sort (LL.cpp:1772) followed by
// an unstable numpy argsort (:41): similarity desc, then template_id asc (Match::operator<, LL.h:233-240), then
// class, y, x ascending.  Exact duplicates (what std::unique, LL.cpp:1773, removes when adjacent) have IoU 1 with
// their twin and die with it, so they never need removing first.
#pragma once
#include <stdint.h>
#include "linemod_b200.h"

struct PostInfo {  // per entry of the selected template sequence
  int32_t class_index, template_id, width, height;
};

struct PostParams {
  const lm_result_header* hdr;  // result block: header + records (all shards' records in fused multi-GPU mode)
  int32_t capacity;
  const PostInfo* info;         // [n_sel], indexed by lm_record.work
  int32_t n_sel;
  double iou_threshold;
  int32_t top_k;                // <= 0: every survivor (bounded by out_capacity)
  lm_match* out;                // device: survivors in pick order
  int32_t out_capacity;
  int32_t* out_counts;          // [0] survivors written, [1] records seen, [2] survivors in total (if all were asked for), [3] coarse candidates
  uint8_t* live;                // [capacity] scratch
};

__device__ __forceinline__ uint32_t lm_float_order(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotone in f
}

struct PostKey {
  unsigned long long a;  // similarity (order-preserving) : 32 | ~template_id : 20 | ~class : 12
  uint32_t b;            // ~y : 16 | ~x : 16
  int32_t idx;
};

__device__ __forceinline__ bool post_better(const PostKey& p, const PostKey& q) {
  if (p.a != q.a) return p.a > q.a;
  if (p.b != q.b) return p.b > q.b;
  return p.idx < q.idx;  // exact twins: the lower slot (either would do)
}

__global__ void __launch_bounds__(1024) k_post_nms(PostParams p) {
  lm_pdl_wait();  // launched with programmatic stream serialization: the records / count of k_refine (k_peer_collect)
                  // are only visible after this returns
  __shared__ PostKey s_best[32];
  __shared__ PostKey s_pick;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = min(p.hdr->count, p.capacity);
  const lm_record* rec = reinterpret_cast<const lm_record*>(p.hdr + 1);
  for (int i = tid; i < n; i += blockDim.x) p.live[i] = 1;
  __syncthreads();
  const int want = p.top_k > 0 ? min(p.top_k, p.out_capacity) : p.out_capacity;
  int picked = 0, survivors = 0;
  for (;;) {
    // block-wide argmax over the live records
    PostKey best;
    best.a = 0ull; best.b = 0u; best.idx = -1;
    for (int i = tid; i < n; i += blockDim.x) {
      if (!p.live[i]) continue;
      const lm_record r = rec[i];
      const int w = min(max(r.work, 0), p.n_sel - 1);
      const PostInfo inf = p.info[w];
      PostKey k;
      k.a = ((unsigned long long)lm_float_order(r.similarity) << 32) |
            ((unsigned long long)(0xFFFFFu - (uint32_t)(inf.template_id & 0xFFFFF)) << 12) |
            (unsigned long long)(0xFFFu - (uint32_t)(inf.class_index & 0xFFF));
      k.b = ((0xFFFFu - (uint32_t)(uint16_t)(r.y + 0x8000)) << 16) | (0xFFFFu - (uint32_t)(uint16_t)(r.x + 0x8000));
      k.idx = i;
      if (best.idx < 0 || post_better(k, best)) best = k;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      PostKey o;
      o.a = __shfl_down_sync(0xffffffffu, best.a, d);
      o.b = __shfl_down_sync(0xffffffffu, best.b, d);
      o.idx = __shfl_down_sync(0xffffffffu, best.idx, d);
      if (o.idx >= 0 && (best.idx < 0 || post_better(o, best))) best = o;
    }
    if (lane == 0) s_best[wid] = best;
    __syncthreads();
    if (wid == 0) {
      best = (lane < (int)(blockDim.x >> 5)) ? s_best[lane] : PostKey{0ull, 0u, -1};
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        PostKey o;
        o.a = __shfl_down_sync(0xffffffffu, best.a, d);
        o.b = __shfl_down_sync(0xffffffffu, best.b, d);
        o.idx = __shfl_down_sync(0xffffffffu, best.idx, d);
        if (o.idx >= 0 && (best.idx < 0 || post_better(o, best))) best = o;
      }
      if (lane == 0) s_pick = best;
    }
    __syncthreads();
    const int pick = s_pick.idx;
    if (pick < 0) break;  // nothing live
    ++survivors;
    const lm_record pr = rec[pick];
    const PostInfo pi = p.info[min(max(pr.work, 0), p.n_sel - 1)];
    if (picked < want) {
      if (tid == 0) {
        lm_match m;
        m.x = pr.x; m.y = pr.y; m.similarity = pr.similarity;
        m.class_index = pi.class_index; m.template_id = pi.template_id;
        p.out[picked] = m;
      }
      ++picked;
    }
    if (p.top_k > 0 && picked >= want) break;
    // suppress: ovr = inter / (area_i + area_j - inter) > thresh, float64, +1 convention (:40, :51-56)
    const double x1 = pr.x, y1 = pr.y, x2 = (double)pr.x + pi.width, y2 = (double)pr.y + pi.height;
    const double area = (x2 - x1 + 1.0) * (y2 - y1 + 1.0);
    for (int i = tid; i < n; i += blockDim.x) {
      if (!p.live[i]) continue;
      if (i == pick) { p.live[i] = 0; continue; }
      const lm_record r = rec[i];
      const PostInfo inf = p.info[min(max(r.work, 0), p.n_sel - 1)];
      const double a1 = r.x, b1 = r.y, a2 = (double)r.x + inf.width, b2 = (double)r.y + inf.height;
      const double w = fmax(0.0, fmin(x2, a2) - fmax(x1, a1) + 1.0);
      const double h = fmax(0.0, fmin(y2, b2) - fmax(y1, b1) + 1.0);
      const double inter = w * h;
      const double ovr = inter / (area + (a2 - a1 + 1.0) * (b2 - b1 + 1.0) - inter);
      if (!(ovr <= p.iou_threshold)) p.live[i] = 0;
    }
    __syncthreads();
  }
  if (tid == 0) {
    p.out_counts[0] = picked;
    p.out_counts[1] = p.hdr->count;
    p.out_counts[2] = p.top_k > 0 ? -1 : survivors;
    p.out_counts[3] = p.hdr->coarse_candidates;  // the host checks it against the candidate list's capacity
  }
}
