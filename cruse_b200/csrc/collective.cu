// collective.cu -- the ONE collective of the path behind the C ABI (SURVEY.md section 8b "flat_allreduce", rows a10 / e):
// g <- sum over ranks of g, then g <- g / world, on the flat fp32 gradient buffer of the training step
// (loss_func/distrib.py:100-116: all_reduce(SUM) per parameter followed by a division by the world size;
// train_base/trainer/base_trainer.py:31 reaches the same result through DistributedDataParallel).
// One ncclAllReduce in place on the caller's stream + one scaling pass; over NVLink 5 / NVSwitch the 12.9 MB of
// config B are latency-, not bandwidth-bound (measured through torch.distributed: 62 us at 2 ranks, 99 us at 8).
//
// NCCL is resolved at RUN time from the libnccl.so.2 the process already holds (torch's own copy; RTLD_NOLOAD first), so
// libcruse_sm100.so has no link-time dependency on it and loads on a box without NCCL -- these entry points then return an error.
#include "common.cuh"
#include <dlfcn.h>
#include <cstring>

namespace cruse {
namespace {

struct NcclUniqueId { char internal[128]; };                       // nccl.h: NCCL_UNIQUE_ID_BYTES = 128
typedef int (*GetUniqueIdFn)(NcclUniqueId*);
typedef int (*CommInitRankFn)(void**, int, NcclUniqueId, int);     // (comm*, nranks, id BY VALUE, rank)
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*CommDestroyFn)(void*);
typedef const char* (*GetErrorStringFn)(int);
constexpr int kNcclFloat32 = 7, kNcclSum = 0;                      // ncclDataType_t / ncclRedOp_t values of nccl.h

struct NcclApi {
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    AllReduceFn all_reduce = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    GetErrorStringFn error_string = nullptr;
    bool ok = false;
};

NcclApi load_nccl() {
    NcclApi a;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy torch.distributed already initialised, if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) return a;
    a.get_unique_id = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
    a.comm_init_rank = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
    a.all_reduce = (AllReduceFn)dlsym(h, "ncclAllReduce");
    a.comm_destroy = (CommDestroyFn)dlsym(h, "ncclCommDestroy");
    a.error_string = (GetErrorStringFn)dlsym(h, "ncclGetErrorString");
    a.ok = a.get_unique_id && a.comm_init_rank && a.all_reduce && a.comm_destroy;
    return a;
}

const NcclApi& nccl() {
    static const NcclApi api = load_nccl();
    return api;
}

int nccl_fail(const char* what, int rc) {
    const NcclApi& a = nccl();
    set_error("%s failed: NCCL error %d (%s)", what, rc, a.error_string ? a.error_string(rc) : "?");
    return -2;
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ buf, long long n, float s) {
    const long long n4 = n >> 2;
    float4* b4 = reinterpret_cast<float4*>(buf);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = b4[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        b4[i] = v;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) buf[i] *= s;
}

}  // namespace
}  // namespace cruse

using namespace cruse;

extern "C" int cruse_nccl_available(void) { return nccl().ok ? 1 : 0; }

extern "C" int cruse_nccl_unique_id(void* id128) {
    CRUSE_CHECK_ARG(id128, "nccl_unique_id: null pointer");
    CRUSE_CHECK_ARG(nccl().ok, "nccl_unique_id: libnccl.so.2 is not available in this process");
    NcclUniqueId id;
    if (int rc = nccl().get_unique_id(&id)) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id128, &id, sizeof(id));
    return 0;
}

extern "C" int cruse_nccl_comm_init(void** comm, int nranks, int rank, const void* id128) {
    CRUSE_CHECK_ARG(comm && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "nccl_comm_init: bad arguments (nranks=%d rank=%d)", nranks, rank);
    CRUSE_CHECK_ARG(nccl().ok, "nccl_comm_init: libnccl.so.2 is not available in this process");
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    *comm = nullptr;
    if (int rc = nccl().comm_init_rank(comm, nranks, id, rank)) return nccl_fail("ncclCommInitRank", rc);
    return 0;
}

extern "C" int cruse_flat_allreduce(void* comm, float* buf, long long n, float scale, void* stream) {
    CRUSE_CHECK_ARG(comm && buf && n > 0, "flat_allreduce: bad arguments");
    CRUSE_CHECK_ARG((reinterpret_cast<uintptr_t>(buf) & 15) == 0, "flat_allreduce: the buffer must be 16-byte aligned");
    CRUSE_CHECK_ARG(nccl().ok, "flat_allreduce: libnccl.so.2 is not available in this process");
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = nccl().all_reduce(buf, buf, (size_t)n, kNcclFloat32, kNcclSum, comm, st)) return nccl_fail("ncclAllReduce", rc);
    if (scale != 1.f) {
        long long blocks = (n / 4 + 255) / 256;
        const long long cap = (long long)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        scale_kernel<<<(unsigned)blocks, 256, 0, st>>>(buf, n, scale);
        CRUSE_LAUNCH_OK();
    }
    return 0;
}

extern "C" int cruse_nccl_comm_destroy(void* comm) {
    CRUSE_CHECK_ARG(comm, "nccl_comm_destroy: null communicator");
    CRUSE_CHECK_ARG(nccl().ok, "nccl_comm_destroy: libnccl.so.2 is not available in this process");
    if (int rc = nccl().comm_destroy(comm)) return nccl_fail("ncclCommDestroy", rc);
    return 0;
}
