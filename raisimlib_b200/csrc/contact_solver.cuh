// Stage D of the fused step: contact::BisectionContactSolver::solve for ONE environment per warp.
//
// Per-contact Gauss-Seidel (Hwangbo, Lee, Hutter, RA-L 2018, section IV; same math as oracle/rbd_oracle.hpp solve_one() /
// step()).  Lane r owns constraint row r: its velocity u_r and impulse lam_r live in registers; a contact update broadcasts the
// six values of contact i by shuffle, evaluates the per-contact rule redundantly on every lane (opening / stick / slip) and each
// lane applies its own row of the Delassus matrix.  The slip branch spreads 32 probes of the (cone surface) x (zero normal
// velocity) curve over the lanes.  Kept OUT OF LINE (one call per sub-step): the solver gets a register allocation of its own,
// so the shared-memory bases stay in registers instead of being re-materialised around every contact update (that cost 25
// instructions per update when this code was inlined into the 72-register kernel body).
//
// Per-contact constant block (CB_WORDS floats, 16-byte aligned, written by stage C):
//   [ a b cc d | e f mu - | Gi0 Gi1 Gi2 Gi3 | Gi4 Gi5 - - ]   G_ii = [a b cc; b d e; cc e f], Gi = its inverse (xx xy xz yy yz zz)
#pragma once
#include <cuda_runtime.h>
#include "../../include/rsb.h"

namespace rsb {

constexpr int CB_WORDS = 16;
constexpr int NREF = 2;              // regula-falsi (Illinois) steps after the 32-section rounds, oracle NREF
constexpr int ACCEL_MAX_RESETS = 2;  // oracle ACCEL_MAX_RESETS

struct GsResult { float lam, resid; int iters, status; };   // status: RSB_SOLVER_* (rsb.h)

// a / b for b > 1e-12 (or a NaN-producing b whose result is discarded): one MUFU.RCP and one multiply, no denormal scaling
__device__ __forceinline__ float fast_div(float a, float b) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return a * r;
}

struct SlipProbe { float g, lx, ly, lz; };
// One probe of the curve {cone surface} x {v_n+ = 0}: lam(theta) = lz (mu cos, mu sin, 1), lz = -c_z / D,
// D = G_zz + mu (G_zx cos + G_zy sin); g = d/dtheta of the energy c.lam + 1/2 lam^T G lam, NaN where the curve does not exist
// (D <= 0), so that every comparison with it is false.
__device__ __forceinline__ SlipProbe slip_probe(float cs, float sn, float a, float b, float cc, float d, float e, float f, float cx, float cy, float cz, float mu) {
  SlipProbe p;
  const float D = fmaf(mu, fmaf(cc, cs, e * sn), f);
  const float lz = fast_div(-cz, D);
  p.lz = lz; p.lx = mu * lz * cs; p.ly = mu * lz * sn;
  const float vx = fmaf(cc, lz, fmaf(b, p.ly, fmaf(a, p.lx, cx)));
  const float vy = fmaf(e, lz, fmaf(d, p.ly, fmaf(b, p.lx, cy)));
  const float g = (vy * cs - vx * sn) * D - mu * (e * cs - cc * sn) * (vx * cs + vy * sn);
  p.g = D > 1e-12f ? g : __int_as_float(0x7fc00000);
  return p;
}
__device__ __forceinline__ float slip_energy(const SlipProbe& p, float a, float b, float cc, float d, float e, float cx, float cy, float cz) {
  const float vx = cx + a * p.lx + b * p.ly + cc * p.lz;
  const float vy = cy + b * p.lx + d * p.ly + e * p.lz;
  return cx * p.lx + cy * p.ly + cz * p.lz + 0.5f * (p.lx * vx + p.ly * vy - p.lz * cz) - 0.5f * (cx * p.lx + cy * p.ly);
}

// Rare outcomes of a 32-probe round, kept out of line: several sign changes (least energy at the left end wins, lowest index on
// ties) and no sign change on the whole circle (least-energy probe).  Returns the lane to take.
__device__ __noinline__ int slip_rank(unsigned cand_mask, bool cand, bool ok, float energy) {
  const bool use = cand_mask ? cand : ok;
  const float fv = use ? energy : 3.0e38f;
  float fm = fv;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fm = fminf(fm, __shfl_xor_sync(0xffffffffu, fm, o));
  const unsigned w = __ballot_sync(0xffffffffu, use && fv == fm);
  return __ffs(w) - 1;
}

// Slip branch of the per-contact rule.  (pc, ps, pv): slip direction of this contact in the previous Gauss-Seidel sweep (valid?)
// in, new direction out.  sec = [cos round 0 | cos round 1 | sin round 0 | sin round 1], SEC_STRIDE floats each.
__device__ __forceinline__ void slip_solve(float a, float b, float cc, float d, float e, float f, float mu, float cx, float cy, float cz, const float* sec, int sec_stride,
                                           int lane, float& pc, float& ps, bool& pv, float& nx, float& ny, float& nz) {
  constexpr unsigned FULLM = 0xffffffffu;
  const float* sec_s = sec + 2 * sec_stride;
  float lo_c = 1.f, lo_s = 0.f, hi_c = 1.f, hi_s = 0.f, glo = 0.f, ghi = 0.f;
  bool have = false;
  const int nxt = (lane + 1) & 31;
  if (pv) {   // local fan: 31 sections of the round-1 table centred on the previous direction (turned back by pi/32)
    const float hc = 0.99518472667219693f, hs = 0.09801714032956060f;
    const float bc = pc * hc + ps * hs, bs = ps * hc - pc * hs;
    const float tc = sec[sec_stride + lane], ts = sec_s[sec_stride + lane];
    const float cs = bc * tc - bs * ts, sn = bs * tc + bc * ts;
    const SlipProbe p = slip_probe(cs, sn, a, b, cc, d, e, f, cx, cy, cz, mu);
    const float g_next = __shfl_sync(FULLM, p.g, nxt);
    const bool cand = (p.g < 0.f) && (g_next >= 0.f) && lane < 31;
    const unsigned m = __ballot_sync(FULLM, cand);
    if (m) {
      int pick = __ffs(m) - 1;
      if (m & (m - 1)) pick = slip_rank(m, cand, true, slip_energy(p, a, b, cc, d, e, cx, cy, cz));
      lo_c = __shfl_sync(FULLM, cs, pick); lo_s = __shfl_sync(FULLM, sn, pick); glo = __shfl_sync(FULLM, p.g, pick);
      hi_c = __shfl_sync(FULLM, cs, pick + 1); hi_s = __shfl_sync(FULLM, sn, pick + 1); ghi = __shfl_sync(FULLM, p.g, pick + 1);
      have = true;
    }
  }
  if (!have) {
    float bc = 1.f, bs = 0.f;
#pragma unroll 1
    for (int r = 0; r < 2; r++) {   // full circle, then its chosen section: bracket 2 pi / 32^(r+1)
      const float tc = sec[r * sec_stride + lane], ts = sec_s[r * sec_stride + lane];
      const float cs = bc * tc - bs * ts, sn = bs * tc + bc * ts;
      const SlipProbe p = slip_probe(cs, sn, a, b, cc, d, e, f, cx, cy, cz, mu);
      float g_next = __shfl_sync(FULLM, p.g, nxt);          // round 0: the direction after probe 31 is probe 0 again
      if (r == 1 && lane == 31) g_next = ghi;              // round 1: it is the upper end of the round-0 bracket (known)
      const bool cand = (p.g < 0.f) && (g_next >= 0.f);
      const unsigned m = __ballot_sync(FULLM, cand);
      if (m == 0u) {
        if (r == 1) break;                                 // keep the round-0 bracket
        // no sign change on the whole circle: least-energy probe; no probe at all: the frictionless normal impulse
        const bool ok = p.g == p.g;
        if (__ballot_sync(FULLM, ok) == 0u) { nx = 0.f; ny = 0.f; nz = fmaxf(0.f, fast_div(-cz, f)); pv = false; return; }
        const int src = slip_rank(0u, false, ok, slip_energy(p, a, b, cc, d, e, cx, cy, cz));
        nx = __shfl_sync(FULLM, p.lx, src); ny = __shfl_sync(FULLM, p.ly, src); nz = __shfl_sync(FULLM, p.lz, src);
        pv = false;
        return;
      }
      int pick = __ffs(m) - 1;
      if (m & (m - 1)) pick = slip_rank(m, cand, true, slip_energy(p, a, b, cc, d, e, cx, cy, cz));
      const int pn = (pick + 1) & 31;
      const float nc = __shfl_sync(FULLM, cs, pn), ns = __shfl_sync(FULLM, sn, pn), ng = __shfl_sync(FULLM, p.g, pn);
      if (!(r == 1 && pick == 31)) { hi_c = nc; hi_s = ns; ghi = ng; }
      lo_c = __shfl_sync(FULLM, cs, pick); lo_s = __shfl_sync(FULLM, sn, pick); glo = __shfl_sync(FULLM, p.g, pick);
      bc = lo_c; bs = lo_s; have = true;
    }
  }
  // regula falsi inside the bracket, Illinois variant (an end kept twice has its value halved); ONE probe per step, evaluated
  // redundantly by every lane (no shuffles).  The last probe is the answer.
  float wlo = glo, whi = ghi;
  int side = 0;
  float cs = lo_c, sn = lo_s;
  nx = 0.f; ny = 0.f; nz = 0.f;   // overwritten by the first step: the curve exists on the whole chord between two valid probes
#pragma unroll
  for (int st = 0; st < NREF; st++) {
    const float tt = fast_div(wlo, wlo - whi);
    float c2 = fmaf(tt, hi_c - lo_c, lo_c), s2 = fmaf(tt, hi_s - lo_s, lo_s);
    const float inv = rsqrtf(c2 * c2 + s2 * s2);
    c2 *= inv; s2 *= inv;
    const SlipProbe p = slip_probe(c2, s2, a, b, cc, d, e, f, cx, cy, cz, mu);
    if (!(p.g == p.g)) break;
    cs = c2; sn = s2; nx = p.lx; ny = p.ly; nz = p.lz;
    if (p.g < 0.f) { lo_c = c2; lo_s = s2; wlo = p.g; if (side == -1) whi *= 0.5f; side = -1; }
    else { hi_c = c2; hi_s = s2; whi = p.g; if (side == 1) wlo *= 0.5f; side = 1; }
  }
  pc = cs; ps = sn; pv = true;
}

// ------------------------------------------------------------------ Anderson acceleration ------
// One step of Anderson acceleration (history of two differences) on the Gauss-Seidel sweep map, oracle step() "accel_m".
// hist = [u0 | x | g1 | f1 | dG | dF] x 32 lanes: u0 = constraint velocities at lambda = 0, x = impulses at the start of this
// sweep, (g1, f1) = output and residual of the previous sweep, (dG, dF) = the difference before that.  lam = this sweep's
// output g; f = g - x.  hc = number of earlier sweeps in the history (0..2), fp = |f|^2 of the previous sweep.
// Cold path (only problems that need more than accel_start - 2 sweeps get here): out of line.
struct AAState { float lam, u, fp; int hc, dropped; };
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __noinline__ AAState anderson_step(float* hist, const float* s_G, int g_stride, int lane, int CR, float lam_c, float u_c, int hc, float fp, int extrapolate) {
  float* h_u0 = hist; float* h_x = hist + 32; float* h_g1 = hist + 64; float* h_f1 = hist + 96; float* h_dg = hist + 128; float* h_df = hist + 160;
  const bool on = lane < CR;
  const float g = on ? lam_c : 0.f;
  const float f = on ? g - h_x[lane] : 0.f;
  const float fn = warp_sum(f * f);
  const float g1 = on ? h_g1[lane] : 0.f, f1 = on ? h_f1[lane] : 0.f;
  AAState o; o.lam = lam_c; o.u = u_c; o.dropped = 0;
  if (hc >= 1 && fn > 4.f * fp) { hc = 0; o.dropped = 1; }   // residual doubled: drop the history, go on from the plain sweep output
  else if (hc >= 1 && extrapolate) {
    const float dFb = f - f1, dGb = g - g1;
    const float dFa = (hc == 2 && on) ? h_df[lane] : 0.f, dGa = (hc == 2 && on) ? h_dg[lane] : 0.f;
    const float a00 = warp_sum(dFa * dFa), a01 = warp_sum(dFa * dFb), r0 = warp_sum(dFa * f);
    const float a11 = warp_sum(dFb * dFb), r1 = warp_sum(dFb * f);
    const float ridge = 1e-10f * (a00 + a11) + 1e-30f;
    float gam0 = 0.f, gam1 = 0.f; bool ok;
    if (hc == 2) {   // 2 x 2 normal equations, elimination without pivoting (SPD + ridge)
      const float p0 = a00 + ridge;
      ok = p0 > 0.f;
      const float m = a01 / p0;
      const float p1 = (a11 + ridge) - m * a01, q1 = r1 - m * r0;
      ok = ok && p1 > 0.f;
      gam1 = q1 / p1; gam0 = (r0 - a01 * gam1) / p0;
    } else {
      const float p1 = a11 + ridge;
      ok = p1 > 0.f;
      gam1 = r1 / p1;
    }
    if (ok) {
      const float xn = g - gam0 * dGa - gam1 * dGb;
      float acc = on ? h_u0[lane] : 0.f;      // u = u0 + G x_next
#pragma unroll 4      // independent shuffles and loads in flight; the sum keeps its order
      for (int b2 = 0; b2 < CR; b2++) {
        const float xb = __shfl_sync(0xffffffffu, xn, b2);
        if (on) acc += s_G[lane * g_stride + b2] * xb;
      }
      o.lam = on ? xn : lam_c; o.u = on ? acc : u_c;
    }
  }
  if (on) {
    if (hc >= 1) { h_dg[lane] = g - g1; h_df[lane] = f - f1; }
    h_g1[lane] = g; h_f1[lane] = f;
  }
  o.hc = min(hc + 1, 2); o.fp = fn;
  return o;
}

// The per-contact rule is cycling on this contact set (typically a joint stop fighting a sticking contact of the same leg):
// continue on G + eps I, eps = stall_reg * mean(diag G) -- constraint-force mixing, only for these problems (oracle step()).
// Updates the Delassus diagonal, the per-contact blocks and their inverses; returns this lane's constraint velocity.
__device__ __noinline__ float regularise(float stall_reg, float* s_G, int g_stride, float* s_cb, int lane, int K, int CR, float lam_c, float u_c) {
  const bool on = lane < CR;
  __syncwarp();   // the sweep that just ended read the per-contact blocks rewritten below
  const float eps = stall_reg * warp_sum(on ? s_G[lane * g_stride + lane] : 0.f) / (float)CR;
  if (on) s_G[lane * g_stride + lane] += eps;
  if (lane < K) {
    float* o = s_cb + CB_WORDS * lane;
    const float a = o[0] + eps, bq = o[1], cc = o[2], d = o[3] + eps, e = o[4], f = o[5] + eps;
    const float c00 = d * f - e * e, c01 = cc * e - bq * f, c02 = bq * e - cc * d;
    const float c11 = a * f - cc * cc, c12 = bq * cc - a * e, c22 = a * d - bq * bq;
    const float id = 1.0f / (a * c00 + bq * c01 + cc * c02);
    o[0] = a; o[3] = d; o[5] = f;
    o[8] = c00 * id; o[9] = c01 * id; o[10] = c02 * id; o[11] = c11 * id; o[12] = c12 * id; o[13] = c22 * id;
  }
  __syncwarp();
  return on ? u_c + eps * lam_c : u_c;
}

// ------------------------------------------------------------------ the solve ------------------
// s_G: Delassus matrix, row stride g_stride (odd: lane-strided row reads are conflict-free); s_cb: per-contact constant blocks;
// u_c: this lane's constraint velocity at lambda = 0 (rows 3K .. 3K+Lm-1 are joint-limit rows).  Returns this lane's impulse.
__device__ __noinline__ GsResult gs_solve(const rsb_params& prm, float* s_G, int g_stride, float* s_cb, float* s_hist, const float* sec, int sec_stride,
                                          int lane, int K, int Lm, float u_c) {
  constexpr unsigned FULLM = 0xffffffffu;
  const int C3 = 3 * K, CR = C3 + Lm;
  const float* Grow = s_G + min(lane, CR - 1) * g_stride;   // lanes past the last row shadow it: branch-free updates, values never used
  float lam_c = 0.f;
  float alpha = prm.alpha_init;
  float sd_c = 1.f, sd_s = 0.f; int sd_v = 0;   // lane i < K: slip direction of contact i in the previous sweep
  float err_ckpt = 3.0e38f;
  int next_ckpt = prm.stall_window;
  int aa_hc = 0, aa_resets = 0; float aa_fp = 0.f;
  bool regularised = false;
  const int aa_first = prm.accel_m > 0 ? prm.accel_start - 2 : 0x7fffffff;   // the history starts two sweeps before the first extrapolation
  if (prm.accel_m > 0) s_hist[lane] = u_c;   // u0
  GsResult res; res.iters = 0; res.resid = 0.f; res.status = RSB_SOLVER_MAXITER;
#pragma unroll 1
  for (int it = 0; it < prm.max_iter; it++) {
    float err = 0.f;
    const bool aa_rec = it + 1 >= aa_first && aa_resets < ACCEL_MAX_RESETS;
    if (aa_rec) s_hist[32 + lane] = lam_c;
#pragma unroll 1
    for (int i = 0; i < K; i++) {
      const int i3 = 3 * i;
      const float ux = __shfl_sync(FULLM, u_c, i3), uy = __shfl_sync(FULLM, u_c, i3 + 1), uz = __shfl_sync(FULLM, u_c, i3 + 2);
      const float lx = __shfl_sync(FULLM, lam_c, i3), ly = __shfl_sync(FULLM, lam_c, i3 + 1), lz = __shfl_sync(FULLM, lam_c, i3 + 2);
      const float4 g0 = *reinterpret_cast<const float4*>(s_cb + CB_WORDS * i);
      const float4 g1 = *reinterpret_cast<const float4*>(s_cb + CB_WORDS * i + 4);
      const float a = g0.x, b = g0.y, cc = g0.z, d = g0.w, e = g1.x, f = g1.y, mu = g1.z;
      // contact velocity without this contact's own impulse
      const float cx = ux - (a * lx + b * ly + cc * lz), cy = uy - (b * lx + d * ly + e * lz), cz = uz - (cc * lx + e * ly + f * lz);
      float nx = 0.f, ny = 0.f, nz = 0.f;
      bool slipped = false;
      if (!(cz > 0.f)) {   // not opening
        const float4 q0 = *reinterpret_cast<const float4*>(s_cb + CB_WORDS * i + 8);
        const float2 q1 = *reinterpret_cast<const float2*>(s_cb + CB_WORDS * i + 12);
        nx = -(q0.x * cx + q0.y * cy + q0.z * cz); ny = -(q0.y * cx + q0.w * cy + q1.x * cz); nz = -(q0.z * cx + q1.x * cy + q1.y * cz);   // stick: -G_ii^-1 c
        if (!(nz >= 0.f && nx * nx + ny * ny <= mu * mu * nz * nz)) {
          float pc = __shfl_sync(FULLM, sd_c, i), ps = __shfl_sync(FULLM, sd_s, i);
          bool pv = __shfl_sync(FULLM, sd_v, i) != 0;
          slip_solve(a, b, cc, d, e, f, mu, cx, cy, cz, sec, sec_stride, lane, pc, ps, pv, nx, ny, nz);
          if (lane == i) { sd_c = pc; sd_s = ps; }
          slipped = pv;
        }
      }
      if (lane == i) sd_v = slipped ? 1 : 0;
      const float dx = alpha * (nx - lx), dy = alpha * (ny - ly), dz = alpha * (nz - lz);
      u_c += Grow[i3] * dx + Grow[i3 + 1] * dy + Grow[i3 + 2] * dz;
      {   // the three rows of contact i take their new impulse: selects, not a divergent branch (ncu: 6 % of the solve sat on its reconvergence)
        const float l0 = lx + dx, l1 = ly + dy, l2 = lz + dz;
        const int rel = lane - i3;
        lam_c = rel == 0 ? l0 : lam_c;
        lam_c = rel == 1 ? l1 : lam_c;
        lam_c = rel == 2 ? l2 : lam_c;
      }
      err = fmaxf(err, fmaxf(fabsf(dx), fmaxf(fabsf(dy), fabsf(dz))));
    }
#pragma unroll 1
    for (int l = 0; l < Lm; l++) {   // joint limits: lam >= 0, complementary to sign * qdot+ - target >= 0
      const int r = C3 + l;
      const float ur = __shfl_sync(FULLM, u_c, r), lr = __shfl_sync(FULLM, lam_c, r);
      const float Grr = s_G[r * g_stride + r];
      const float ln = fmaxf(0.f, -(ur - Grr * lr) / Grr);
      const float dl = alpha * (ln - lr);
      u_c += Grow[r] * dl;
      if (lane == r) lam_c = lr + dl;
      err = fmaxf(err, fabsf(dl));
    }
    res.iters = it + 1; res.resid = err;
    alpha = fmaxf(prm.alpha_min, alpha * prm.alpha_decay);
    if (err < prm.threshold) { res.status = regularised ? RSB_SOLVER_CONVERGED_COMPLIANT : RSB_SOLVER_CONVERGED; break; }
    if (it + 1 == next_ckpt) {      // stagnation check (see rsb_params.stall_window)
      if (it + 1 >= 2 * prm.stall_window && err > prm.stall_ratio * err_ckpt) {
        if (regularised || !(prm.stall_reg > 0.f)) { res.status = RSB_SOLVER_STALLED; break; }
        u_c = regularise(prm.stall_reg, s_G, g_stride, s_cb, lane, K, CR, lam_c, u_c);   // first stall: go on with a slightly compliant contact set
        regularised = true; aa_hc = 0; aa_resets = 0;
        err_ckpt = 3.0e38f; next_ckpt = it + 1 + prm.stall_window;
        continue;
      }
      err_ckpt = err; next_ckpt += prm.stall_window;
    }
    if (aa_rec && it + 1 < prm.max_iter) {   // after the stagnation check, never on the last sweep: the loop always ends on a projected sweep's impulses
      __syncwarp();
      const AAState st = anderson_step(s_hist, s_G, g_stride, lane, CR, lam_c, u_c, aa_hc, aa_fp, it + 1 >= prm.accel_start ? 1 : 0);
      lam_c = st.lam; u_c = st.u; aa_hc = st.hc; aa_fp = st.fp; aa_resets += st.dropped;
    }
  }
  res.lam = lam_c;
  return res;
}

}  // namespace rsb
