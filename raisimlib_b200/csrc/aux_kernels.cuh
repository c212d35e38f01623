// Observation row of the RaisimGym ANYmal locomotion task ([RECALL]
// raisimGymTorch envs/rsg_anymal/Environment.hpp updateObservation()):
//   [ z | R^T e_z (3) | joint q | R^T v_base (3) | R^T w_base (3) | joint rates ]
// One warp per environment, coalesced row reads and writes; fixed-base models get [q | qdot].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rsb.h"

namespace rsb {

// obs_rot_column / obs_dot3: see step_kernel.cuh (shared with the fused observation write)
__global__ void rsb_observe_kernel(const float* __restrict__ gc, const float* __restrict__ gv, int gc_stride, int gv_stride, int nq, int nv,
                                   int floating, int num_envs, float* __restrict__ obs, int ob_dim) {
  const int lane = threadIdx.x & 31;
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (env >= num_envs) return;
  const float* q = gc + (size_t)env * gc_stride;
  const float* v = gv + (size_t)env * gv_stride;
  float* o = obs + (size_t)env * ob_dim;
  if (!floating) {
    for (int i = lane; i < nq; i += 32) o[i] = q[i];
    for (int i = lane; i < nv; i += 32) o[nq + i] = v[i];
    return;
  }
  float qw = q[3], qx = q[4], qy = q[5], qz = q[6];
  float inv = 1.0f / sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  qw *= inv; qx *= inv; qy *= inv; qz *= inv;
  const int nj = nq - 7;
  if (lane == 0) o[0] = q[2];
  if (lane < 3) {
    float c0, c1, c2;
    obs_rot_column(qw, qx, qy, qz, lane, c0, c1, c2);
    o[1 + lane] = c2;
    o[4 + nj + lane] = obs_dot3(c0, c1, c2, v[0], v[1], v[2]);
    o[7 + nj + lane] = obs_dot3(c0, c1, c2, v[3], v[4], v[5]);
  }
  for (int i = lane; i < nj; i += 32) { o[4 + i] = q[7 + i]; o[10 + nj + i] = v[6 + i]; }
}

// external wrench rows for rsb_step_kernel (StepArgs::ext): [body | F(3) | T(3) | point in body frame(3) | pad(2)]
__global__ void rsb_ext_pack_kernel(float* rows, int body, const float* __restrict__ force, const float* __restrict__ torque, float px, float py, float pz,
                                    int count) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  float* r = rows + (size_t)e * EXT_WORDS;
  r[0] = __int_as_float(body);
  for (int k = 0; k < 3; k++) { r[1 + k] = force ? force[(size_t)e * 3 + k] : 0.f; r[4 + k] = torque ? torque[(size_t)e * 3 + k] : 0.f; }
  r[7] = px; r[8] = py; r[9] = pz; r[10] = 0.f; r[11] = 0.f;
}

}  // namespace rsb

// ---- RaisimGym ANYmal locomotion task on the device ([RECALL] raisimGymTorch
//      envs/rsg_anymal/Environment.hpp + VectorizedEnvironment.hpp::perAgentStep) -----------------------
namespace rsb {

struct GymConfig {
  const float* gc_init;      // [nq]
  const float* gv_init;      // [nv]
  const float* action_mean;  // [nq - 7]
  const float* action_std;   // [nq - 7]
  uint32_t foot_mask;        // bit b set: contacts on body b do not terminate the episode
  float torque_coeff, forward_vel_coeff, terminal_reward;
};

// ENVIRONMENT::step() first half: pTarget.tail(nJoints) = action * actionStd + actionMean
__global__ void rsb_gym_action_kernel(const float* __restrict__ action, GymConfig cfg, int nq, int gc_stride, int num_envs, float* __restrict__ ptarget) {
  const int nj = nq - 7;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= num_envs * nj) return;
  const int env = idx / nj, j = idx % nj;
  ptarget[(size_t)env * gc_stride + 7 + j] = action[idx] * cfg.action_std[j] + cfg.action_mean[j];
}

// ENVIRONMENT::reset() for every environment
__global__ void rsb_gym_reset_kernel(float* gc, float* gv, float* ptarget, float* vtarget, GymConfig cfg, int gc_stride, int gv_stride, int nq, int nv, int num_envs) {
  const int lane = threadIdx.x & 31;
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (env >= num_envs) return;
  for (int i = lane; i < nq; i += 32) { gc[(size_t)env * gc_stride + i] = cfg.gc_init[i]; ptarget[(size_t)env * gc_stride + i] = cfg.gc_init[i]; }
  for (int i = lane; i < nv; i += 32) { gv[(size_t)env * gv_stride + i] = cfg.gv_init[i]; vtarget[(size_t)env * gv_stride + i] = 0.f; }
}

// second half, one warp per environment: reward from the post-step state, isTerminalState(), reset(), observe()
__global__ void rsb_gym_post_kernel(float* gc, float* gv, const float* __restrict__ tau_applied, float* __restrict__ ptarget,
                                    const int* __restrict__ ncontacts, const rsb_contact* __restrict__ contacts, GymConfig cfg, int gc_stride,
                                    int gv_stride, int nq, int nv, int num_envs, float* __restrict__ obs, int ob_dim, float* __restrict__ reward,
                                    unsigned char* __restrict__ done) {
  const int lane = threadIdx.x & 31;
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (env >= num_envs) return;
  float* q = gc + (size_t)env * gc_stride;
  float* v = gv + (size_t)env * gv_stride;
  const float* ta = tau_applied + (size_t)env * gv_stride;
  float qw = q[3], qx = q[4], qy = q[5], qz = q[6];
  float inv = 1.0f / sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  qw *= inv; qx *= inv; qy *= inv; qz *= inv;
  // first column of R -> body-frame x velocity = R(:,0) . v
  const float r00 = 1.f - 2.f * (qy * qy + qz * qz), r10 = 2.f * (qx * qy + qw * qz), r20 = 2.f * (qx * qz - qw * qy);
  const float vbx = r00 * v[0] + r10 * v[1] + r20 * v[2];
  float t2 = 0.f;
  for (int i = lane; i < nv; i += 32) t2 += ta[i] * ta[i];
  for (int o = 16; o > 0; o >>= 1) t2 += __shfl_xor_sync(0xffffffffu, t2, o);
  const int K = ncontacts[env];
  bool bad = false;
  if (lane < K) { const int b = contacts[(size_t)env * RSB_KMAX + lane].local_body; bad = ((cfg.foot_mask >> b) & 1u) == 0u; }
  const bool term = __any_sync(0xffffffffu, bad);
  float r = cfg.torque_coeff * t2 + cfg.forward_vel_coeff * fminf(4.0f, vbx);
  if (term) {
    r += cfg.terminal_reward;
    for (int i = lane; i < nq; i += 32) { q[i] = cfg.gc_init[i]; ptarget[(size_t)env * gc_stride + i] = cfg.gc_init[i]; }
    for (int i = lane; i < nv; i += 32) v[i] = cfg.gv_init[i];
  }
  __syncwarp();
  if (lane == 0) { reward[env] = r; done[env] = term ? 1 : 0; }
  // observation row of the (possibly reset) state
  qw = q[3]; qx = q[4]; qy = q[5]; qz = q[6];
  inv = 1.0f / sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  qw *= inv; qx *= inv; qy *= inv; qz *= inv;
  float R[9] = {1.f - 2.f * (qy * qy + qz * qz), 2.f * (qx * qy - qw * qz), 2.f * (qx * qz + qw * qy),
                2.f * (qx * qy + qw * qz), 1.f - 2.f * (qx * qx + qz * qz), 2.f * (qy * qz - qw * qx),
                2.f * (qx * qz - qw * qy), 2.f * (qy * qz + qw * qx), 1.f - 2.f * (qx * qx + qy * qy)};
  float* o = obs + (size_t)env * ob_dim;
  const int nj = nq - 7;
  if (lane == 0) o[0] = q[2];
  if (lane < 3) {
    o[1 + lane] = R[6 + lane];
    o[4 + nj + lane] = R[0 + lane] * v[0] + R[3 + lane] * v[1] + R[6 + lane] * v[2];
    o[7 + nj + lane] = R[0 + lane] * v[3] + R[3 + lane] * v[4] + R[6 + lane] * v[5];
  }
  for (int i = lane; i < nj; i += 32) { o[4 + i] = q[7 + i]; o[10 + nj + i] = v[6 + i]; }
}

}  // namespace rsb
