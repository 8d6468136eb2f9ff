// GreedyGeodesicFollowerImpl::nextBestPrimAlong (GreedyFollower.cpp:83-140) for N agents at once, with the
// kinematics and the selection on the device.  The reference evaluates each primitive
// `[LEFT]*k + [FORWARD]` / `[RIGHT]*k + [FORWARD]` by bouncing C++ -> Python -> C++ through a MoveFn
// (GreedyFollower.h:54) and running try_step, find_path and distance_to_closest_obstacle one at a time
// (GreedyFollower.cpp:46-81).  Here:
//   k_follower_targets  thread per agent: the 2 * nSteps headings (default_controls.py: rotate about +Y,
//                       move along local -Z; float64 quaternions, the operation order of
//                       nav/greedy_follower.py's numpy restatement) -> forward targets of every primitive
//   (try_step, find_path, closest_obstacle over all N * 2 * nSteps primitives: the batched kernels)
//   k_follower_select   thread per agent: computeReward (GreedyFollower.cpp:62-81) per primitive in its
//                       float arithmetic and the sequential selection (first strictly better reward,
//                       early exit above 0.99 after a LEFT / RIGHT pair, GreedyFollower.cpp:99-137)
#pragma once
#include <cuda_runtime.h>
#include "hbn_math.h"

namespace hbn {

struct FollowerParams {
  float goalDist, forwardAmount;
  double sinHalf, cosHalf;  // of turn_amount / 2
  int nSteps;               // headings per side: angle = 0, t, 2t, ... < pi (accumulated in f32 by the host)
};

enum { kFolError = -2, kFolStop = -1, kFolNone = -3 };  // else (k << 1) | side, side 0 = LEFT, 1 = RIGHT

__device__ __forceinline__ void folQuatMul(const double* a, const double* b, double* o) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by - ax * bz + ay * bw + az * bx;
  o[2] = aw * bz + ax * by - ay * bx + az * bw;
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
}
__device__ __forceinline__ void folCross(const double* u, const double* v, double* o) {
  o[0] = u[1] * v[2] - u[2] * v[1];
  o[1] = u[2] * v[0] - u[0] * v[2];
  o[2] = u[0] * v[1] - u[1] * v[0];
}
// pos + rotate(q, (0, 0, -forward)): v + 2 * cross(u, cross(u, v) + w * v)
__device__ __forceinline__ void folForwardTarget(const double* q, const double* pos, double fwd, double* o) {
  const double v[3] = {0.0, 0.0, -fwd};
  double c1[3], inner[3], c2[3];
  folCross(q, v, c1);
  for (int k = 0; k < 3; ++k) inner[k] = c1[k] + q[3] * v[k];
  folCross(q, inner, c2);
  for (int k = 0; k < 3; ++k) o[k] = pos[k] + (v[k] + 2.0 * c2[k]);
}
__device__ __forceinline__ void folTurn(double* q, double s, double c) {
  const double y[4] = {0.0, s, 0.0, c};
  double t[4];
  folQuatMul(q, y, t);
  const double n = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + t[3] * t[3]);
  for (int k = 0; k < 4; ++k) q[k] = t[k] / n;
}

// starts / targets / ends: [n, 2 * nSteps, 3] f32; candidate 2k = k LEFT turns, 2k + 1 = k RIGHT turns
__global__ void __launch_bounds__(128) k_follower_targets(const double* __restrict__ rots, const double* __restrict__ poss,
                                                          const float* __restrict__ goals, int64_t n, FollowerParams p,
                                                          float* __restrict__ pos32, float* __restrict__ starts,
                                                          float* __restrict__ targets, float* __restrict__ ends) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double l[4], r[4], pos[3];
  for (int k = 0; k < 4; ++k) l[k] = r[k] = rots[4 * i + k];
  for (int k = 0; k < 3; ++k) pos[k] = poss[3 * i + k];
  const float s32[3] = {static_cast<float>(pos[0]), static_cast<float>(pos[1]), static_cast<float>(pos[2])};
  const float g32[3] = {goals[3 * i], goals[3 * i + 1], goals[3 * i + 2]};
  for (int k = 0; k < 3; ++k) pos32[3 * i + k] = s32[k];
  const double fwd = static_cast<double>(p.forwardAmount);
  for (int k = 0; k < p.nSteps; ++k) {
    double tl[3], tr[3];
    folForwardTarget(l, pos, fwd, tl);
    folForwardTarget(r, pos, fwd, tr);
    const size_t c = (static_cast<size_t>(i) * 2 * p.nSteps + 2 * k) * 3;
    for (int j = 0; j < 3; ++j) {
      targets[c + j] = static_cast<float>(tl[j]);
      targets[c + 3 + j] = static_cast<float>(tr[j]);
      starts[c + j] = starts[c + 3 + j] = s32[j];
      ends[c + j] = ends[c + 3 + j] = g32[j];
    }
    folTurn(l, p.sinHalf, p.cosHalf);
    folTurn(r, -p.sinHalf, p.cosHalf);
  }
}

__global__ void __launch_bounds__(128) k_follower_select(const float* __restrict__ geo0, const float* __restrict__ starts,
                                                         const float* __restrict__ targets, const float* __restrict__ filt,
                                                         const float* __restrict__ geoAfter, const float* __restrict__ obsAfter,
                                                         int64_t n, FollowerParams p, int32_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float geo = geo0[i];
  if (geo == infF() || geo != geo) {
    out[i] = kFolError;
    return;
  }
  if (geo < p.goalDist) {
    out[i] = kFolStop;
    return;
  }
  float bestReward = -0.25f;  // -collisionCost_: "we are just constantly colliding"
  int best = -1;
  for (int c = 0; c < 2 * p.nSteps; ++c) {
    const size_t o = (static_cast<size_t>(i) * 2 * p.nSteps + c) * 3;
    float before = 0.f, after = 0.f;
    {
      const float d0 = targets[o] - starts[o], d1 = targets[o + 1] - starts[o + 1], d2 = targets[o + 2] - starts[o + 2];
      before = d0 * d0 + d1 * d1;
      before += d2 * d2;
      const float e0 = filt[o] - starts[o], e1 = filt[o + 1] - starts[o + 1], e2 = filt[o + 2] - starts[o + 2];
      after = e0 * e0 + e1 * e1;
      after += e2 * e2;
    }
    const bool collided = (after + 1e-5f) < before;  // object_controls.py: EPS
    float penalty = -0.0125f * static_cast<float>(c >> 1);
    penalty = penalty - (collided ? 0.25f : 0.0f);
    penalty = penalty - (obsAfter[o / 3] < 0.2f ? 0.05f : 0.0f);
    const float reward = (geo - geoAfter[o / 3]) / p.forwardAmount + penalty;
    if (reward > bestReward) {
      bestReward = reward;
      best = c;
    }
    if ((c & 1) && bestReward > 0.99f) break;  // GreedyFollower.cpp:131-135
  }
  out[i] = best < 0 ? kFolNone : best;
}

}  // namespace hbn
