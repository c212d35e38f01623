// rsb_comm_*: the path's one collective for C++ hosts -- all-gather of the observation rows over NCCL
// (NVLink 5 / NVSwitch), one rsb_batch per GPU inside ONE process (bench.py does the same with
// torch.distributed, one process per GPU).  NCCL is bound at run time (dlopen) so that librsb.so carries no
// link-time dependency and a process that already loaded NCCL (PyTorch) shares that copy.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/rsb.h"

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool load() {
    if (lib) return true;
    const char* names[] = {getenv("RSB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { error = "NCCL not found (set RSB_NCCL_LIB to libnccl.so.2)"; return false; }
    auto sym = [&](const char* s) { void* p = dlsym(lib, s); if (!p) error = std::string("NCCL symbol missing: ") + s; return p; };
    CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    AllGather = (decltype(AllGather))sym("ncclAllGather");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd && GetErrorString;
  }
};
NcclApi g_nccl;
}  // namespace

struct rsb_comm {
  std::vector<rsb_batch*> batches;
  std::vector<int> devices;
  std::vector<ncclComm_t> comms;
  std::vector<cudaStream_t> streams;
  std::vector<float*> obs_local;   // per device: [n_local][ob_dim]
  int ob_dim = 0, n_local = 0;
};

// set by batch.cu
extern "C" int rsb_internal_batch_info(rsb_batch* b, int* device, void** stream, int* num_envs);
extern "C" void rsb_internal_set_error(const char* msg);

static int cfail(const std::string& m) { rsb_internal_set_error(m.c_str()); return RSB_ERR_CUDA; }

extern "C" {

int rsb_comm_init(rsb_batch** batches, int ndev, rsb_comm** out) {
  if (!batches || ndev < 1 || !out) { rsb_internal_set_error("bad arguments to rsb_comm_init"); return RSB_ERR_INVALID; }
  if (!g_nccl.load()) return cfail(g_nccl.error);
  rsb_comm* c = new rsb_comm;
  c->batches.assign(batches, batches + ndev);
  c->devices.resize(ndev); c->streams.resize(ndev); c->comms.resize(ndev); c->obs_local.assign(ndev, nullptr);
  c->ob_dim = rsb_batch_ob_dim(batches[0]);
  for (int i = 0; i < ndev; i++) {
    int n = 0; void* s = nullptr;
    if (rsb_internal_batch_info(batches[i], &c->devices[i], &s, &n) != RSB_OK) { delete c; return RSB_ERR_INVALID; }
    c->streams[i] = (cudaStream_t)s;
    if (i == 0) c->n_local = n;
    if (n != c->n_local || rsb_batch_ob_dim(batches[i]) != c->ob_dim) { delete c; rsb_internal_set_error("rsb_comm_init: every batch must hold the same number of environments of the same model"); return RSB_ERR_INVALID; }
  }
  ncclResult_t r = g_nccl.CommInitAll(c->comms.data(), ndev, c->devices.data());
  if (r != ncclSuccess) { std::string m = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r); delete c; return cfail(m); }
  for (int i = 0; i < ndev; i++) {
    cudaSetDevice(c->devices[i]);
    if (cudaMalloc((void**)&c->obs_local[i], (size_t)c->n_local * c->ob_dim * 4) != cudaSuccess) { rsb_comm_destroy(c); return cfail("rsb_comm_init: cudaMalloc failed"); }
  }
  *out = c;
  return RSB_OK;
}

// every device: observation rows of its own environments, then ncclAllGather into obs_all[dev] ([ndev * n_local][ob_dim],
// device memory on that GPU), all inside one NCCL group on the batches' own streams (asynchronous).
int rsb_comm_allgather_obs(rsb_comm* c, float* const* obs_all) {
  if (!c || !obs_all) { rsb_internal_set_error("null argument to rsb_comm_allgather_obs"); return RSB_ERR_INVALID; }
  const int ndev = (int)c->batches.size();
  for (int i = 0; i < ndev; i++) {
    int rc = rsb_batch_observe(c->batches[i], c->obs_local[i], 0, c->n_local, RSB_DEVICE);
    if (rc != RSB_OK) return rc;
  }
  ncclResult_t r = g_nccl.GroupStart();
  for (int i = 0; i < ndev && r == ncclSuccess; i++)
    r = g_nccl.AllGather(c->obs_local[i], obs_all[i], (size_t)c->n_local * c->ob_dim, ncclFloat, c->comms[i], c->streams[i]);
  ncclResult_t r2 = g_nccl.GroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess) return cfail(std::string("ncclAllGather: ") + g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
  return RSB_OK;
}

void rsb_comm_destroy(rsb_comm* c) {
  if (!c) return;
  for (size_t i = 0; i < c->comms.size(); i++) {
    cudaSetDevice(c->devices[i]);
    if (c->obs_local[i]) cudaFree(c->obs_local[i]);
    if (c->comms[i] && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comms[i]);
  }
  delete c;
}

}  // extern "C"
