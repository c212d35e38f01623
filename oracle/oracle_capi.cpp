// ORACLE C API -- test infrastructure, NOT product code (see rbd_oracle.hpp header).
// Builds liboracle.so: a double and a float instance of the restated World::integrate(),
// OpenMP over environments the way RaisimGym's VectorizedEnvironment does
// (`#pragma omp parallel for` over envs, SURVEY.md 3.1 [RECALL]).
#include "counting_scalar.hpp"
#include "rbd_oracle.hpp"
#include <omp.h>
#include <memory>

using namespace orc;

struct OrcDebug {       // every pointer optional; sized for n_envs; filled from the LAST step
  double* M;            // [n][nv*nv]   mass matrix (no PD augmentation)
  double* h;            // [n][nv]      bias force incl. gravity
  double* R;            // [n][nb*9]
  double* p;            // [n][nb*3]
  int* ncontacts;       // [n]
  int* c_pt;            // [n][KMAX]
  int* c_body;          // [n][KMAX]
  int* c_pair;          // [n][KMAX]
  double* c_pos;        // [n][KMAX*3]
  double* c_normal;     // [n][KMAX*3]
  double* c_depth;      // [n][KMAX]
  double* c_lambda;     // [n][KMAX*3]  contact-frame impulse (t1,t2,n)
  int* iters;           // [n]
  double* G;            // [n][RMAX^2]  Delassus matrix of the constraint rows: contacts, then joint limits (row stride RMAX)
  double* u0;           // [n][RMAX]    free constraint velocity minus target
  int ext_body;         // IN external wrench for this call (all sub-steps): body index, < 0 = none
  const double* ext_force;   // [n][3] world, nullable
  const double* ext_torque;  // [n][3] world, nullable
  double ext_point[3];  // application point in the body frame
  int* warm_pt;         // [n][KMAX]   IN/OUT contact cache (candidate-point ids, -1 = empty); null = cold start
  double* warm_imp;     // [n][KMAX*3] IN/OUT world-frame impulses of the cache
  double* tau_applied;  // [n][nv]     generalized force applied over the last step
  int* nlimits;         // [n]         active joint-limit rows
  int* lim_dof;         // [n][LMAX]   their dofs (-1 = none)
  double* lim_lambda;   // [n][LMAX]   their impulses
  double* resid;        // [n]         largest impulse update of the last sweep (solver residual at exit)
  int* status;          // [n]         how the solve ended (RSB_SOLVER_*)
};

struct Handle {
  int precision;
  std::unique_ptr<Sim<double>> d;
  std::unique_ptr<Sim<float>> f;
  std::unique_ptr<Sim<Cnt>> c;      // precision 2: FLOP-counting build (single-threaded use)
};

long long g_counts[4] = {0, 0, 0, 0};   // per-contact rule outcomes summed over all run() calls (statistics for tools/, tests)
// counters of the FLOP-counting build (precision 2): add, mul, div, sqrt, trig, cmp
extern "C" void orc_get_flops(long long* out, int reset) {
  orc::FlopCounters& c = orc::flop_counters();
  out[0] = c.add; out[1] = c.mul; out[2] = c.div; out[3] = c.sqrt; out[4] = c.trig; out[5] = c.cmp;
  if (reset) c = orc::FlopCounters();
}
extern "C" void orc_get_counts(long long* out, int reset) { for (int k = 0; k < 4; k++) { out[k] = g_counts[k]; if (reset) g_counts[k] = 0; } }

template <typename T>
static void run(Sim<T>& sim, int n_envs, int n_steps, double* gc, double* gv, const double* tau, const double* pt, const double* vt,
                const double* kp, const double* kd, int nthreads, OrcDebug* dbg) {
  const int nq = sim.nq, nv = sim.nv, nb = sim.nb;
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
  {
    Workspace<T> ws;
    sim.init_ws(ws);
    std::vector<T> q(nq), v(nv), tf(nv), ptt(nq), vtt(nv), kpp(nv), kdd(nv);
#pragma omp for schedule(static)
    for (int e = 0; e < n_envs; e++) {
      for (int i = 0; i < nq; i++) q[i] = T(gc[(size_t)e * nq + i]);
      for (int i = 0; i < nv; i++) v[i] = T(gv[(size_t)e * nv + i]);
      if (tau) for (int i = 0; i < nv; i++) tf[i] = T(tau[(size_t)e * nv + i]);
      if (pt) for (int i = 0; i < nq; i++) ptt[i] = T(pt[(size_t)e * nq + i]);
      if (vt) for (int i = 0; i < nv; i++) vtt[i] = T(vt[(size_t)e * nv + i]);
      if (kp) for (int i = 0; i < nv; i++) kpp[i] = T(kp[i]);
      if (kd) for (int i = 0; i < nv; i++) kdd[i] = T(kd[i]);
      for (int k = 0; k < KMAX; k++) {
        ws.prev_pt[k] = (dbg && dbg->warm_pt) ? dbg->warm_pt[(size_t)e * KMAX + k] : -1;
        if (dbg && dbg->warm_imp) ws.prev_imp[k] = {T(dbg->warm_imp[((size_t)e * KMAX + k) * 3]), T(dbg->warm_imp[((size_t)e * KMAX + k) * 3 + 1]), T(dbg->warm_imp[((size_t)e * KMAX + k) * 3 + 2])};
      }
      ws.hm_offset = (sim.ter.type == 2 && !sim.ter.env_map.empty()) ? sim.ter.env_map[(size_t)e % sim.ter.env_map.size()] * sim.ter.xs * sim.ter.ys : 0;
      ws.ext_body = -1;
      if (dbg && dbg->ext_body >= 0) {
        ws.ext_body = dbg->ext_body;
        ws.ext_f = dbg->ext_force ? V3<T>{T(dbg->ext_force[3 * (size_t)e]), T(dbg->ext_force[3 * (size_t)e + 1]), T(dbg->ext_force[3 * (size_t)e + 2])} : V3<T>{0, 0, 0};
        ws.ext_t = dbg->ext_torque ? V3<T>{T(dbg->ext_torque[3 * (size_t)e]), T(dbg->ext_torque[3 * (size_t)e + 1]), T(dbg->ext_torque[3 * (size_t)e + 2])} : V3<T>{0, 0, 0};
        ws.ext_pos = {T(dbg->ext_point[0]), T(dbg->ext_point[1]), T(dbg->ext_point[2])};
      }
      for (int s = 0; s < n_steps; s++)
        sim.step(q.data(), v.data(), tau ? tf.data() : nullptr, pt ? ptt.data() : nullptr, vt ? vtt.data() : nullptr,
                 kp ? kpp.data() : nullptr, kd ? kdd.data() : nullptr, ws);
      for (int i = 0; i < nq; i++) gc[(size_t)e * nq + i] = double(q[i]);
      for (int i = 0; i < nv; i++) gv[(size_t)e * nv + i] = double(v[i]);
      if (dbg && dbg->warm_pt) for (int k = 0; k < KMAX; k++) {
        dbg->warm_pt[(size_t)e * KMAX + k] = ws.prev_pt[k];
        if (dbg->warm_imp) { dbg->warm_imp[((size_t)e * KMAX + k) * 3] = double(ws.prev_imp[k].x); dbg->warm_imp[((size_t)e * KMAX + k) * 3 + 1] = double(ws.prev_imp[k].y); dbg->warm_imp[((size_t)e * KMAX + k) * 3 + 2] = double(ws.prev_imp[k].z); }
      }
      if (dbg) {
        if (dbg->M) for (int i = 0; i < nv * nv; i++) dbg->M[(size_t)e * nv * nv + i] = double(ws.M[i]);
        if (dbg->h) for (int i = 0; i < nv; i++) dbg->h[(size_t)e * nv + i] = double(ws.h[i]);
        if (dbg->tau_applied) for (int i = 0; i < nv; i++) dbg->tau_applied[(size_t)e * nv + i] = double(ws.tau_applied[i]);
        if (dbg->R) for (int i = 0; i < nb; i++) for (int k = 0; k < 9; k++) dbg->R[((size_t)e * nb + i) * 9 + k] = double(ws.R[i].m[k]);
        if (dbg->p) for (int i = 0; i < nb; i++) { dbg->p[((size_t)e * nb + i) * 3] = double(ws.p[i].x); dbg->p[((size_t)e * nb + i) * 3 + 1] = double(ws.p[i].y); dbg->p[((size_t)e * nb + i) * 3 + 2] = double(ws.p[i].z); }
        int K = int(ws.contacts.size());
        if (dbg->ncontacts) dbg->ncontacts[e] = K;
        if (dbg->iters) dbg->iters[e] = ws.iters;
        if (dbg->resid) dbg->resid[e] = double(ws.resid);
        if (dbg->status) dbg->status[e] = ws.status;
        const int Crows = 3 * K + int(ws.limits.size());
        if (dbg->G) for (int a = 0; a < Crows; a++) for (int b2 = 0; b2 < Crows; b2++) dbg->G[(size_t)e * RMAX * RMAX + a * RMAX + b2] = double(ws.G[a * Crows + b2]);
        if (dbg->u0) for (int a = 0; a < Crows; a++) dbg->u0[(size_t)e * RMAX + a] = double(ws.u0[a]);
        if (dbg->nlimits) dbg->nlimits[e] = int(ws.limits.size());
        if (dbg->lim_dof) for (int l = 0; l < LMAX; l++) { dbg->lim_dof[(size_t)e * LMAX + l] = l < (int)ws.limits.size() ? ws.limits[l].dof : -1; if (dbg->lim_lambda) dbg->lim_lambda[(size_t)e * LMAX + l] = l < (int)ws.limits.size() ? double(ws.limits[l].lam) : 0.0; }
        for (int k = 0; k < KMAX; k++) {
          size_t o = (size_t)e * KMAX + k;
          bool on = k < K;
          if (dbg->c_pt) dbg->c_pt[o] = on ? ws.contacts[k].pt : -1;
          if (dbg->c_body) dbg->c_body[o] = on ? ws.contacts[k].body : -1;
          if (dbg->c_pair) dbg->c_pair[o] = on ? ws.contacts[k].pair : -1;
          if (dbg->c_depth) dbg->c_depth[o] = on ? double(ws.contacts[k].depth) : 0.0;
          if (dbg->c_pos) { dbg->c_pos[3 * o] = on ? double(ws.contacts[k].pos.x) : 0; dbg->c_pos[3 * o + 1] = on ? double(ws.contacts[k].pos.y) : 0; dbg->c_pos[3 * o + 2] = on ? double(ws.contacts[k].pos.z) : 0; }
          if (dbg->c_normal) { dbg->c_normal[3 * o] = on ? double(ws.contacts[k].n.x) : 0; dbg->c_normal[3 * o + 1] = on ? double(ws.contacts[k].n.y) : 0; dbg->c_normal[3 * o + 2] = on ? double(ws.contacts[k].n.z) : 0; }
          if (dbg->c_lambda) { dbg->c_lambda[3 * o] = on ? double(ws.contacts[k].lam.x) : 0; dbg->c_lambda[3 * o + 1] = on ? double(ws.contacts[k].lam.y) : 0; dbg->c_lambda[3 * o + 2] = on ? double(ws.contacts[k].lam.z) : 0; }
        }
      }
    }
#pragma omp critical
    for (int k = 0; k < 4; k++) g_counts[k] += ws.counts[k];
  }
}

extern "C" {

void* orc_create(const ModelDesc* d, int precision) {
  Handle* h = new Handle;
  h->precision = precision;
  if (precision == 0) h->d.reset(new Sim<double>(*d)); else if (precision == 1) h->f.reset(new Sim<float>(*d)); else h->c.reset(new Sim<Cnt>(*d));
  return h;
}
void orc_destroy(void* hv) { delete static_cast<Handle*>(hv); }

int orc_kmax() { return KMAX; }

// p = {dt, gx, gy, gz, erp, alpha_init, alpha_min, alpha_decay, max_iter, threshold, mu, restitution, rest_threshold, stall_window, stall_ratio, warm_start, slip_bisect, joint_limits, slip_local, accel_m, accel_start, stall_reg}
void orc_set_params(void* hv, const double* p) {
  Handle* h = static_cast<Handle*>(hv);
  Params prm;
  prm.dt = p[0]; prm.gravity[0] = p[1]; prm.gravity[1] = p[2]; prm.gravity[2] = p[3]; prm.erp = p[4];
  prm.alpha_init = p[5]; prm.alpha_min = p[6]; prm.alpha_decay = p[7]; prm.max_iter = int(p[8]); prm.threshold = p[9];
  prm.mu = p[10]; prm.restitution = p[11]; prm.rest_threshold = p[12]; prm.stall_window = int(p[13]); prm.stall_ratio = p[14]; prm.warm_start = int(p[15]); prm.slip_bisect = int(p[16]); prm.joint_limits = int(p[17]); prm.slip_local = int(p[18]); prm.accel_m = int(p[19]); prm.accel_start = int(p[20]); prm.stall_reg = p[21];
  if (h->d) h->d->prm = prm; else if (h->f) h->f->prm = prm; else h->c->prm = prm;
}

void orc_set_ground(void* hv, double z) {
  Handle* h = static_cast<Handle*>(hv);
  Terrain t; t.type = 1; t.ground_z = z;
  if (h->d) h->d->set_terrain(t); else if (h->f) h->f->set_terrain(t); else h->c->set_terrain(t);
}

void orc_set_heightmap(void* hv, int xs, int ys, double x_size, double y_size, double cx, double cy, const double* heights) {
  Handle* h = static_cast<Handle*>(hv);
  Terrain t; t.type = 2; t.xs = xs; t.ys = ys; t.x_size = x_size; t.y_size = y_size; t.cx = cx; t.cy = cy;
  t.h.assign(heights, heights + (size_t)xs * ys);
  if (h->d) h->d->set_terrain(t); else if (h->f) h->f->set_terrain(t); else h->c->set_terrain(t);
}

// terrain atlas: `count` same-sized height maps back to back, env_map[n_envs] picks one per environment
void orc_set_heightmaps(void* hv, int count, int xs, int ys, double x_size, double y_size, double cx, double cy, const double* heights,
                        const int* env_map, int n_envs) {
  Handle* h = static_cast<Handle*>(hv);
  Terrain t; t.type = 2; t.xs = xs; t.ys = ys; t.x_size = x_size; t.y_size = y_size; t.cx = cx; t.cy = cy; t.count = count;
  t.h.assign(heights, heights + (size_t)count * xs * ys);
  t.env_map.assign(env_map, env_map + n_envs);
  if (h->d) h->d->set_terrain(t); else if (h->f) h->f->set_terrain(t); else h->c->set_terrain(t);
}

void orc_clear_terrain(void* hv) {
  Handle* h = static_cast<Handle*>(hv);
  Terrain t;
  if (h->d) h->d->set_terrain(t); else if (h->f) h->f->set_terrain(t); else h->c->set_terrain(t);
}

// n_steps of World::integrate() for n_envs environments; gc [n][nq], gv [n][nv] updated in place.
// tau_ff [n][nv], ptarget [n][nq], vtarget [n][nv] per env (nullable); kp, kd [nv] shared (nullable).
int orc_step(void* hv, int n_envs, int n_steps, double* gc, double* gv, const double* tau_ff, const double* ptarget,
             const double* vtarget, const double* kp, const double* kd, int nthreads, OrcDebug* dbg) {
  Handle* h = static_cast<Handle*>(hv);
  if (h->d) run(*h->d, n_envs, n_steps, gc, gv, tau_ff, ptarget, vtarget, kp, kd, nthreads, dbg);
  else if (h->c) run(*h->c, n_envs, n_steps, gc, gv, tau_ff, ptarget, vtarget, kp, kd, 1, dbg);
  else run(*h->f, n_envs, n_steps, gc, gv, tau_ff, ptarget, vtarget, kp, kd, nthreads, dbg);
  return 0;
}

// isolated per-contact solve, for the solver unit tests: G 3x3 row-major (t1,t2,n), c[3], mu -> lam[3]
void orc_solve_one(void* hv, const double* G, const double* c, double mu, double* lam) {
  Handle* h = static_cast<Handle*>(hv);
  if (h->d) {
    V3<double> l; h->d->solve_one(G, V3<double>{c[0], c[1], c[2]}, mu, l);
    lam[0] = l.x; lam[1] = l.y; lam[2] = l.z;
  } else {
    float Gf[9]; for (int i = 0; i < 9; i++) Gf[i] = float(G[i]);
    V3<float> l; h->f->solve_one(Gf, V3<float>{float(c[0]), float(c[1]), float(c[2])}, float(mu), l);
    lam[0] = l.x; lam[1] = l.y; lam[2] = l.z;
  }
}

// per-candidate-point friction override (< 0: default material)
void orc_set_point_mu(void* hv, const double* mu) {
  Handle* h = static_cast<Handle*>(hv);
  if (h->d) for (int i = 0; i < h->d->npts; i++) h->d->pt_mu[i] = mu[i];
  else for (int i = 0; i < h->f->npts; i++) h->f->pt_mu[i] = float(mu[i]);
}

int orc_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
