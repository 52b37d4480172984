// Gradient all-reduce of the data-parallel train step (the reference is single-device, main.py --gpu G).
//
// One NCCL communicator per handle, created from a unique id the caller broadcasts (vc_comm_unique_id on rank 0 ->
// vc_comm_init on every rank). libnccl is bound at run time with dlopen: inside a PyTorch process this resolves to
// the copy torch already loaded, a standalone caller gets the system library, and the CPU-only test box needs none.
//
// The collective is bucketed in the order the backward pass produces gradients (vocabulary projection first, CNN
// last): Model::grad_ready(ranges) records an event on the compute stream, a high-priority communication stream
// waits for it and sums the ranges in place over NVLink / NVSwitch while the compute stream keeps going; apply()
// joins the two streams before the global norm and Adam read the reduced buffer.
#include <dlfcn.h>
#include <cstdlib>
#include <mutex>
#include "model.h"

namespace vc {

namespace {
typedef int ncclResult;
typedef void* ncclCommPtr;
struct NcclId {
  char internal[128];
};
struct Nccl {
  void* lib = nullptr;
  ncclResult (*GetVersion)(int*) = nullptr;
  ncclResult (*GetUniqueId)(NcclId*) = nullptr;
  ncclResult (*CommInitRank)(ncclCommPtr*, int, NcclId, int) = nullptr;
  ncclResult (*CommDestroy)(ncclCommPtr) = nullptr;
  ncclResult (*AllReduce)(const void*, void*, size_t, int, int, ncclCommPtr, cudaStream_t) = nullptr;
  ncclResult (*GroupStart)() = nullptr;
  ncclResult (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult) = nullptr;
  std::string err;
};
constexpr int kNcclFloat32 = 7, kNcclBfloat16 = 9, kNcclSum = 0;  // nccl.h enum values (stable since NCCL 2.10)

// bf16 transport (VC_GRAD_BF16=1): fp32 gradients are rounded to bf16 for the wire and widened back afterwards, both on
// the communication stream. Vectorised by 8; ranges are multiples of 64 floats.
__global__ void k_grad_pack(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
    uint4 o;
    o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w); o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
}
__global__ void k_grad_unpack(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(src)[i];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    const float2 p0 = __bfloat1622float2(h[0]), p1 = __bfloat1622float2(h[1]), p2 = __bfloat1622float2(h[2]), p3 = __bfloat1622float2(h[3]);
    reinterpret_cast<float4*>(dst)[2 * i] = make_float4(p0.x, p0.y, p1.x, p1.y);
    reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(p2.x, p2.y, p3.x, p3.y);
  }
}

const Nccl* nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("VC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (nm == nullptr || nm[0] == 0) continue;
      n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (n.lib) break;
    }
    if (!n.lib) {
      n.err = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "?");
      return;
    }
#define VC_SYM(field, name)                                            \
  *(void**)(&n.field) = dlsym(n.lib, name);                            \
  if (!n.field) n.err = std::string("libnccl lacks ") + name;
    VC_SYM(GetVersion, "ncclGetVersion")
    VC_SYM(GetUniqueId, "ncclGetUniqueId")
    VC_SYM(CommInitRank, "ncclCommInitRank")
    VC_SYM(CommDestroy, "ncclCommDestroy")
    VC_SYM(AllReduce, "ncclAllReduce")
    VC_SYM(GroupStart, "ncclGroupStart")
    VC_SYM(GroupEnd, "ncclGroupEnd")
    VC_SYM(GetErrorString, "ncclGetErrorString")
#undef VC_SYM
  });
  return &n;
}

#define VC_NCCL(n, expr)                                                                                          \
  do {                                                                                                            \
    ncclResult _r = (expr);                                                                                       \
    if (_r != 0) return set_error(VC_E_NCCL, "%s failed: %s", #expr, (n)->GetErrorString ? (n)->GetErrorString(_r) : "?"); \
  } while (0)
}  // namespace

struct Comm {
  ncclCommPtr comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  static constexpr int kEvents = 32;
  cudaEvent_t ready[kEvents] = {};
  cudaEvent_t done = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr;  // timing of the first / last collective of a step (vc_comm_stats)
  int next_event = 0;
  bool pending = false, timed = false;
  __nv_bfloat16* wire = nullptr;  // bf16 transport buffer (same element offsets as the gradient buffer); null = fp32 transport
  int mode = 1;  // 0: no reduction (measurement only), 1: bucketed + overlapped, 2: one all-reduce behind the backward pass
  long long bytes_step = 0, calls_step = 0;
};

int comm_unique_id(void* id128) {
  const Nccl* n = nccl();
  if (!n->err.empty()) return set_error(VC_E_NCCL, "%s", n->err.c_str());
  NcclId id;
  VC_NCCL(n, n->GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return VC_OK;
}

int Model::comm_init(const void* id128, int rank, int world) {
  if (comm != nullptr) return set_error(VC_E_STATE, "vc_comm_init: this handle already has a communicator");
  if (world < 1 || rank < 0 || rank >= world) return set_error(VC_E_ARG, "vc_comm_init: rank %d outside world %d", rank, world);
  const Nccl* n = nccl();
  if (!n->err.empty()) return set_error(VC_E_NCCL, "%s", n->err.c_str());
  VC_CUDA(cudaSetDevice(device));
  Comm* c = new Comm();
  c->rank = rank;
  c->world = world;
  NcclId id;
  memcpy(id.internal, id128, 128);
  ncclResult r = n->CommInitRank(&c->comm, world, id, rank);
  if (r != 0) {
    delete c;
    return set_error(VC_E_NCCL, "ncclCommInitRank(rank %d of %d) failed: %s", rank, world, n->GetErrorString(r));
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = highest priority
  VC_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
  for (auto& e : c->ready) VC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  VC_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
  VC_CUDA(cudaEventCreate(&c->t0));
  VC_CUDA(cudaEventCreate(&c->t1));
  if (const char* e = getenv("VC_GRAD_BF16")) {
    if (e[0] == '1') VC_CUDA(cudaMalloc((void**)&c->wire, (size_t)(cfg.fine_tune ? n_total : n_adam + 64) * sizeof(__nv_bfloat16)));
  }
  comm = c;
  return VC_OK;
}

void Model::comm_release() {
  if (comm == nullptr) return;
  cudaSetDevice(device);
  cudaDeviceSynchronize();
  const Nccl* n = nccl();
  if (comm->comm && n->CommDestroy) n->CommDestroy(comm->comm);
  for (auto& e : comm->ready)
    if (e) cudaEventDestroy(e);
  if (comm->done) cudaEventDestroy(comm->done);
  if (comm->t0) cudaEventDestroy(comm->t0);
  if (comm->t1) cudaEventDestroy(comm->t1);
  if (comm->stream) cudaStreamDestroy(comm->stream);
  if (comm->wire) cudaFree(comm->wire);
  delete comm;
  comm = nullptr;
}

int Model::comm_world() const { return comm ? comm->world : 1; }
bool Model::comm_bf16() const { return comm != nullptr && comm->wire != nullptr; }

// Gradients in [off, off + count) of the flat buffer (several ranges per call = one bucket) are final on stream `s`:
// sum them over the ranks on the communication stream. No-op without a communicator or with world == 1.
int Model::grad_ready(const int64_t (*ranges)[2], int n_ranges, cudaStream_t s) {
  if (comm == nullptr || comm->mode != 1) return VC_OK;
  return comm_reduce(ranges, n_ranges, s);
}

int Model::comm_reduce(const int64_t (*ranges)[2], int n_ranges, cudaStream_t s) {
  // a one-rank communicator still goes through NCCL (a copy onto itself): the single-GPU test suite exercises this path
  if (comm == nullptr || n_ranges <= 0) return VC_OK;
  const Nccl* n = nccl();
  Comm& c = *comm;
  cudaEvent_t ev = c.ready[c.next_event];
  c.next_event = (c.next_event + 1) % Comm::kEvents;
  VC_CUDA(cudaEventRecord(ev, s));
  VC_CUDA(cudaStreamWaitEvent(c.stream, ev, 0));
  if (!c.pending) {
    c.bytes_step = 0;
    c.calls_step = 0;
    VC_CUDA(cudaEventRecord(c.t0, c.stream));
    c.timed = true;
  }
  // the 64-float tail (squared norms, O(1e5) under the AG prior) always travels in fp32
  auto wired = [&](int i) { return c.wire != nullptr && ranges[i][1] > 64; };
  for (int i = 0; i < n_ranges; ++i)
    if (wired(i)) {
      const long long n8 = ranges[i][1] / 8;
      k_grad_pack<<<(int)std::min<long long>((n8 + 255) / 256, num_sms() * 4), 256, 0, c.stream>>>(Gf + ranges[i][0], c.wire + ranges[i][0], n8);
    }
  if (n_ranges > 1) VC_NCCL(n, n->GroupStart());
  for (int i = 0; i < n_ranges; ++i) {
    const int64_t off = ranges[i][0], cnt = ranges[i][1];
    if (cnt <= 0) continue;
    ncclResult r = wired(i) ? n->AllReduce(c.wire + off, c.wire + off, (size_t)cnt, kNcclBfloat16, kNcclSum, c.comm, c.stream)
                            : n->AllReduce(Gf + off, Gf + off, (size_t)cnt, kNcclFloat32, kNcclSum, c.comm, c.stream);
    if (r != 0) {
      if (n_ranges > 1) n->GroupEnd();
      return set_error(VC_E_NCCL, "ncclAllReduce(%lld floats) failed: %s", (long long)cnt, n->GetErrorString(r));
    }
    c.bytes_step += cnt * (wired(i) ? 2 : 4);
  }
  if (n_ranges > 1) VC_NCCL(n, n->GroupEnd());
  for (int i = 0; i < n_ranges; ++i)
    if (wired(i)) {
      const long long n8 = ranges[i][1] / 8;
      k_grad_unpack<<<(int)std::min<long long>((n8 + 255) / 256, num_sms() * 4), 256, 0, c.stream>>>(c.wire + ranges[i][0], Gf + ranges[i][0], n8);
    }
  VC_CUDA(cudaGetLastError());
  c.calls_step += 1;
  c.pending = true;
  return VC_OK;
}

// Contiguous span covering parameters [first, last] of the flat buffer (they must be adjacent in the layout).
int Model::grad_ready_params(std::initializer_list<int> ids, cudaStream_t s) {
  if (comm == nullptr || comm->mode != 1) return VC_OK;
  int64_t r[16][2];
  int n = 0;
  for (int pi : ids) {
    if (pi == -2 && n < 16) {  // the 64-float tail behind the clipped-Adam region (embedding-slice squared norms, Q4)
      r[n][0] = n_adam;
      r[n][1] = 64;
      ++n;
      continue;
    }
    if (pi < 0) continue;
    const ParamInfo& p = params[pi];
    if (p.parent >= 0 || p.region == 1) continue;
    const int64_t off = p.offset, cnt = (p.count + 63) / 64 * 64;
    if (n > 0 && r[n - 1][0] + r[n - 1][1] == off) {
      r[n - 1][1] += cnt;
    } else if (n < 16) {
      r[n][0] = off;
      r[n][1] = cnt;
      ++n;
    }
  }
  return grad_ready(r, n, s);
}

// The optimiser's stream waits for every outstanding bucket.
int Model::comm_join(cudaStream_t s) {
  if (comm != nullptr && comm->mode == 2 && !comm->pending) {
    const int64_t r[1][2] = {{0, cfg.fine_tune ? n_total : n_adam + 64}};
    VC_TRY(comm_reduce(r, 1, s));
  }
  if (comm == nullptr || !comm->pending) return VC_OK;
  VC_CUDA(cudaEventRecord(comm->t1, comm->stream));
  VC_CUDA(cudaEventRecord(comm->done, comm->stream));
  VC_CUDA(cudaStreamWaitEvent(s, comm->done, 0));
  comm->pending = false;
  return VC_OK;
}

// Un-overlapped form (vc_allreduce_gradients): the whole flat buffer in one call behind the backward pass.
int Model::comm_allreduce_all(cudaStream_t s) {
  if (comm == nullptr) return set_error(VC_E_STATE, "vc_allreduce_gradients: no communicator (call vc_comm_init)");
  const int64_t r[1][2] = {{0, cfg.fine_tune ? n_total : n_adam + 64}};
  VC_TRY(comm_reduce(r, 1, s));
  return comm_join(s);
}

int Model::comm_set_mode(int mode) {
  if (comm == nullptr) return set_error(VC_E_STATE, "no communicator");
  if (mode < 0 || mode > 2) return set_error(VC_E_ARG, "vc_comm_set_mode: mode must be 0, 1 or 2");
  comm->mode = mode;
  return VC_OK;
}

int Model::comm_stats(float* ms, long long* bytes, int* calls) {
  if (comm == nullptr) return set_error(VC_E_STATE, "no communicator");
  VC_CUDA(cudaStreamSynchronize(comm->stream));
  float t = 0.f;
  if (comm->timed && cudaEventElapsedTime(&t, comm->t0, comm->t1) != cudaSuccess) t = 0.f;
  if (ms) *ms = t;
  if (bytes) *bytes = comm->bytes_step;
  if (calls) *calls = (int)comm->calls_step;
  return VC_OK;
}

}  // namespace vc
