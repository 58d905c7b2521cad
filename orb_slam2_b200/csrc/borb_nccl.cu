// NCCL inside the C ABI: the one collective the path has — the packed vocabulary broadcast once at start-up
// (reference src/System.cc:65 loads ORBvoc.txt on every process; here one rank parses it and NVLink carries the 48 MB blob).
// A C++ Tracking host needs no torch: libnccl.so.2 is resolved at run time with dlopen, so libborb.so has no link-time
// dependency on NCCL and loads on machines without it.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "borb_match.h"

using namespace borb;

namespace {
typedef int nccl_result;
typedef void* nccl_comm;
struct nccl_uid { char internal[128]; };                    // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
constexpr int NCCL_UINT8 = 1;                               // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1
struct NcclApi {
    void* h = nullptr;
    nccl_result (*GetUniqueId)(nccl_uid*) = nullptr;
    nccl_result (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
    nccl_result (*CommDestroy)(nccl_comm) = nullptr;
    nccl_result (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(nccl_result) = nullptr;
};
NcclApi* nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!api.h) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
        api.Broadcast = (decltype(api.Broadcast))dlsym(api.h, "ncclBroadcast");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
    });
    return (api.h && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast) ? &api : nullptr;
}
borb_status nccl_fail(NcclApi* a, nccl_result r, const char* what) {
    set_error("%s failed: %s", what, (a && a->GetErrorString) ? a->GetErrorString(r) : "NCCL error");
    return BORB_ERR_CUDA;
}
}  // namespace

extern "C" {

borb_status borb_nccl_unique_id(uint8_t* id128) {
    if (!id128) { set_error("null argument"); return BORB_ERR_INVALID_ARG; }
    NcclApi* a = nccl();
    if (!a) { set_error("libnccl.so.2 not found"); return BORB_ERR_UNSUPPORTED; }
    nccl_uid u;
    const nccl_result r = a->GetUniqueId(&u);
    if (r != 0) return nccl_fail(a, r, "ncclGetUniqueId");
    std::memcpy(id128, u.internal, 128);
    return BORB_OK;
}

borb_status borb_nccl_comm_create(const uint8_t* id128, int world_size, int rank, int device, void** comm_out) {
    if (!id128 || !comm_out || world_size < 1 || rank < 0 || rank >= world_size) { set_error("bad arguments"); return BORB_ERR_INVALID_ARG; }
    NcclApi* a = nccl();
    if (!a) { set_error("libnccl.so.2 not found"); return BORB_ERR_UNSUPPORTED; }
    BORB_CUDA(cudaSetDevice(device));
    nccl_uid u;
    std::memcpy(u.internal, id128, 128);
    nccl_comm c = nullptr;
    const nccl_result r = a->CommInitRank(&c, world_size, u, rank);
    if (r != 0) return nccl_fail(a, r, "ncclCommInitRank");
    *comm_out = c;
    return BORB_OK;
}

borb_status borb_nccl_comm_destroy(void* comm) {
    NcclApi* a = nccl();
    if (!a || !comm) return BORB_OK;
    a->CommDestroy((nccl_comm)comm);
    return BORB_OK;
}

borb_status borb_voc_broadcast(borb_voc* root_voc, void* nccl_comm_, int root, int rank, int device, borb_voc** out) {
    if (!nccl_comm_ || !out || (rank == root && !root_voc)) { set_error("bad arguments"); return BORB_ERR_INVALID_ARG; }
    NcclApi* a = nccl();
    if (!a) { set_error("libnccl.so.2 not found"); return BORB_ERR_UNSUPPORTED; }
    *out = nullptr;
    BORB_CUDA(cudaSetDevice(device));
    cudaStream_t s = nullptr;
    BORB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    // 1. the blob size (8 bytes through the device: NCCL moves device buffers)
    unsigned long long* d_size = nullptr;
    BORB_CUDA(cudaMalloc(&d_size, 8));
    void* blob = nullptr;
    size_t bytes = 0;
    if (rank == root) {
        borb_status st = borb_voc_blob(root_voc, &blob, &bytes);
        if (st != BORB_OK) return st;
        const unsigned long long b = bytes;
        BORB_CUDA(cudaMemcpyAsync(d_size, &b, 8, cudaMemcpyHostToDevice, s));
    }
    nccl_result r = a->Broadcast(d_size, d_size, 8, NCCL_UINT8, root, (nccl_comm)nccl_comm_, s);
    if (r != 0) return nccl_fail(a, r, "ncclBroadcast(size)");
    unsigned long long b = 0;
    BORB_CUDA(cudaMemcpyAsync(&b, d_size, 8, cudaMemcpyDeviceToHost, s));
    BORB_CUDA(cudaStreamSynchronize(s));
    bytes = (size_t)b;
    // 2. the blob itself, straight into the receiver's HBM
    if (rank != root) BORB_CUDA(cudaMalloc(&blob, bytes));
    r = a->Broadcast(blob, blob, bytes, NCCL_UINT8, root, (nccl_comm)nccl_comm_, s);
    if (r != 0) return nccl_fail(a, r, "ncclBroadcast(blob)");
    BORB_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_size);
    cudaStreamDestroy(s);
    if (rank == root) { *out = root_voc; return BORB_OK; }
    borb_status st = borb_voc_from_blob(blob, bytes, device, out);
    if (st != BORB_OK) { cudaFree(blob); return st; }
    borb_voc_adopt_ownership(*out);            // the receiver allocated the blob: it frees it with the vocabulary
    return BORB_OK;
}

}  // extern "C"
