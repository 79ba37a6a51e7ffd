// rx_mgpu.cu -- rxc_mgpu_*: delivery of finished bands / frames to rank 0 (the reference's serial compose
// of tile buffers into `pixels`, src/rasterizer.rs:560-579, is what this replaces across GPUs).
//
// One process and one rxc_ctx per GPU.  Rank 0 owns the delivery buffer; every other rank maps it into its own address
// space (cudaIpc, i.e. NVLink / NVSwitch peer access) and k_raster's tile write-back stores straight into the mapping:
// the "gather" is the 128-bit stores of the raster kernel itself, overlapped tile by tile with the rendering, and no copy
// or collective follows it.  Completion is a flag per rank in rank 0's memory (system-scope release store after the
// raster kernel / acquire spin on rank 0's stream); the reverse flag lets rank 0 hand a buffer back.  Where a peer
// mapping cannot be had (no P2P between the two GPUs), ranks render into a local staging buffer and the regions travel
// as ONE ncclGroup of send/recv pairs on a second stream.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, the copy already in the process when the host is PyTorch): the
// library has no link-time dependency on it, and a single-GPU host never loads it.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rx_internal.h"

namespace {

// ---- the slice of NCCL's C API that is used (nccl.h, 2.x: stable enum values) -------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclUint8 = 1, kNcclInt32 = 2, kNcclMin = 3 };
struct Nccl {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};

Nccl* load_nccl(std::string* why) {
    static Nccl n;
    static bool tried = false;
    static std::string err;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        // the copy the process already holds (PyTorch's) first: two NCCL instances in one process work, one is better
        for (const char* nm : names) if (!n.so) n.so = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        for (const char* nm : names) if (!n.so) n.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (!n.so) {
            err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
        } else {
#define RX_SYM(field, name) *(void**)(&n.field) = dlsym(n.so, name); if (!n.field) err = std::string("libnccl has no ") + name
            RX_SYM(GetUniqueId, "ncclGetUniqueId"); RX_SYM(CommInitRank, "ncclCommInitRank"); RX_SYM(CommDestroy, "ncclCommDestroy");
            RX_SYM(GetErrorString, "ncclGetErrorString"); RX_SYM(Broadcast, "ncclBroadcast"); RX_SYM(AllReduce, "ncclAllReduce");
            RX_SYM(Send, "ncclSend"); RX_SYM(Recv, "ncclRecv"); RX_SYM(GroupStart, "ncclGroupStart"); RX_SYM(GroupEnd, "ncclGroupEnd");
#undef RX_SYM
        }
    }
    if (!err.empty()) { if (why) *why = err; return nullptr; }
    return &n;
}

// ---- control block at the end of rank 0's delivery buffer ------------------------------------------------------------
#define RX_MGPU_MAX_RANKS 64
struct MgpuCtl {
    uint32_t arrive[RX_MGPU_MAX_RANKS];   // arrive[r] = number of deliveries rank r has completed (written by rank r over the peer mapping)
    uint32_t release;                     // number of releases rank 0 has issued
    uint32_t timeouts;                    // a wait gave up (a rank died): the data of that step is incomplete
    uint32_t pad[62];
};
static_assert(sizeof(MgpuCtl) == 512, "control block layout");

#define RX_MGPU_TIMEOUT_NS 20000000000ull   // 20 s: a wait never hangs the GPU for good when a peer process has died

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The raster kernel that wrote this rank's pixels into the peer mapping has completed (stream order): its stores are
// performed at system scope before the flag is.
__global__ void k_mgpu_signal(uint32_t* flag, uint32_t seq) {
    __threadfence_system();
    st_release_sys(flag, seq);
}

// One lane per awaited flag: spins until flag[i] >= seq (wrap-safe), gives up after RX_MGPU_TIMEOUT_NS.
__global__ void k_mgpu_wait(const uint32_t* flags, uint32_t first, uint32_t n, uint32_t seq, uint32_t* timeouts) {
    const unsigned long long t0 = global_ns();
    for (uint32_t i = first + threadIdx.x; i < first + n; i += blockDim.x) {
        while ((int32_t)(ld_acquire_sys(flags + i) - seq) < 0) {
            __nanosleep(200);
            if (global_ns() - t0 > RX_MGPU_TIMEOUT_NS) { atomicAdd(timeouts, 1u); break; }
        }
    }
    __threadfence_system();
}

}  // namespace

struct RxMgpu {
    Nccl* nccl = nullptr;
    ncclComm_t comm = nullptr;
    uint32_t rank = 0, world = 1;
    uint32_t mode = RXC_MGPU_LOCAL;
    bool force_nccl = false;
    // delivery buffer
    uint8_t* base = nullptr;        // rank 0: its allocation; other ranks: the peer mapping (PEER) or the local staging (NCCL)
    uint64_t bytes = 0;             // payload bytes (the control block follows)
    bool base_is_ipc = false;
    MgpuCtl* ctl = nullptr;         // in rank 0's allocation (local or through the mapping); nullptr in NCCL mode on ranks > 0
    uint32_t seq = 0, rseq = 0;     // deliveries / releases issued so far
    // NCCL mode
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_rendered = nullptr, ev_moved = nullptr;
    bool moved_pending = false;
    uint8_t* pack = nullptr; uint64_t pack_cap = 0;   // contiguous copies of pitched regions
    uint32_t* d_word = nullptr;     // small device scratch for the collectives of the set-up
};

namespace {

#define CKC(ctx, call)                                                                                      \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) return rxi_fail(ctx, RXC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define CKN(ctx, m, call)                                                                                   \
    do {                                                                                                    \
        ncclResult_t r__ = (call);                                                                          \
        if (r__ != 0) return rxi_fail(ctx, RXC_ERR_CUDA, std::string(#call) + ": " + (m)->nccl->GetErrorString(r__)); \
    } while (0)

RxMgpu* state(rxc_ctx* ctx) { return ctx ? *rxi_mgpu_slot(ctx) : nullptr; }

void free_target(rxc_ctx* ctx, RxMgpu* m) {
    cudaSetDevice(rxi_device(ctx));
    cudaStreamSynchronize(rxi_stream(ctx));
    if (m->comm_stream) cudaStreamSynchronize(m->comm_stream);
    if (m->base) {
        if (m->base_is_ipc) cudaIpcCloseMemHandle(m->base); else cudaFree(m->base);
    }
    m->base = nullptr; m->bytes = 0; m->ctl = nullptr; m->base_is_ipc = false; m->seq = 0; m->rseq = 0; m->moved_pending = false;
}

// Collective teardown of the delivery buffer: the importers close their mappings before rank 0 frees the memory.
int32_t free_target_ordered(rxc_ctx* ctx, RxMgpu* m) {
    if (!m->base) return RXC_OK;
    if (m->world > 1 && m->comm) {
        cudaStream_t st = rxi_stream(ctx);
        if (m->rank != 0) free_target(ctx, m);
        CKN(ctx, m, m->nccl->AllReduce(m->d_word + 48, m->d_word + 48, 1, kNcclInt32, kNcclMin, m->comm, st));
        CKC(ctx, cudaStreamSynchronize(st));
    }
    free_target(ctx, m);
    return RXC_OK;
}

template <class F>
int32_t guarded(rxc_ctx* ctx, F&& f) noexcept {
    try {
        return f();
    } catch (const std::bad_alloc&) {
        if (ctx) rxi_fail(ctx, RXC_ERR_OOM, "out of host memory");
        return RXC_ERR_OOM;
    } catch (...) {
        if (ctx) rxi_fail(ctx, RXC_ERR_INVALID, "internal error");
        return RXC_ERR_INVALID;
    }
}

}  // namespace

void rxi_mgpu_destroy(rxc_ctx* ctx) {
    RxMgpu* m = state(ctx);
    if (!m) return;
    free_target(ctx, m);
    if (m->pack) cudaFree(m->pack);
    if (m->d_word) cudaFree(m->d_word);
    if (m->ev_rendered) cudaEventDestroy(m->ev_rendered);
    if (m->ev_moved) cudaEventDestroy(m->ev_moved);
    if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
    if (m->comm && m->nccl) m->nccl->CommDestroy(m->comm);
    delete m;
    *rxi_mgpu_slot(ctx) = nullptr;
}

extern "C" {

int32_t rxc_mgpu_unique_id(uint8_t* id) {
    return guarded(nullptr, [&]() -> int32_t {
    if (!id) return RXC_ERR_INVALID;
    Nccl* n = load_nccl(nullptr);
    if (!n) return RXC_ERR_UNSUPPORTED;
    ncclUniqueId u;
    if (n->GetUniqueId(&u) != 0) return RXC_ERR_CUDA;
    static_assert(sizeof(u) == RXC_MGPU_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &u, sizeof(u));
    return RXC_OK;
    });
}

int32_t rxc_mgpu_init(rxc_ctx* ctx, const uint8_t* id, uint32_t rank, uint32_t world) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    if (world == 0 || rank >= world || world > RX_MGPU_MAX_RANKS) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_init: bad rank / world size");
    if (world > 1 && !id) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_init: the unique id is required when world > 1");
    if (state(ctx)) rxi_mgpu_destroy(ctx);
    CKC(ctx, cudaSetDevice(rxi_device(ctx)));
    RxMgpu* m = new RxMgpu();
    *rxi_mgpu_slot(ctx) = m;
    m->rank = rank; m->world = world;
    if (const char* e = getenv("RXC_MGPU_FORCE_NCCL")) m->force_nccl = atoi(e) != 0;
    if (world > 1) {
        std::string why;
        m->nccl = load_nccl(&why);
        if (!m->nccl) { rxi_mgpu_destroy(ctx); return rxi_fail(ctx, RXC_ERR_UNSUPPORTED, "rxc_mgpu_init: " + why); }
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        ncclResult_t r = m->nccl->CommInitRank(&m->comm, (int)world, u, (int)rank);
        if (r != 0) { std::string s = m->nccl->GetErrorString(r); m->comm = nullptr; rxi_mgpu_destroy(ctx); return rxi_fail(ctx, RXC_ERR_CUDA, "ncclCommInitRank: " + s); }
        CKC(ctx, cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking));
        CKC(ctx, cudaEventCreateWithFlags(&m->ev_rendered, cudaEventDisableTiming));
        CKC(ctx, cudaEventCreateWithFlags(&m->ev_moved, cudaEventDisableTiming));
    }
    CKC(ctx, cudaMalloc((void**)&m->d_word, 256));
    return RXC_OK;
    });
}

int32_t rxc_mgpu_shutdown(rxc_ctx* ctx) {
    return guarded(ctx, [&]() -> int32_t {
    if (!ctx) return RXC_ERR_INVALID;
    if (RxMgpu* m = state(ctx)) {
        cudaSetDevice(rxi_device(ctx));
        const int32_t st = free_target_ordered(ctx, m);
        if (st != RXC_OK) return st;
    }
    rxi_mgpu_destroy(ctx);
    return RXC_OK;
    });
}

int32_t rxc_mgpu_target(rxc_ctx* ctx, uint64_t bytes, void** rank0_ptr, uint32_t* mode_out) {
    return guarded(ctx, [&]() -> int32_t {
    RxMgpu* m = state(ctx);
    if (!m) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_init has not been called");
    if (bytes == 0) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_target: empty buffer");
    CKC(ctx, cudaSetDevice(rxi_device(ctx)));
    cudaStream_t st = rxi_stream(ctx);
    { const int32_t fs = free_target_ordered(ctx, m); if (fs != RXC_OK) return fs; }
    const uint64_t payload = (bytes + 511) & ~(uint64_t)511;
    const uint64_t total = payload + sizeof(MgpuCtl);
    if (rank0_ptr) *rank0_ptr = nullptr;
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    int32_t ok = 1;
    if (m->rank == 0) {
        cudaError_t e = cudaMalloc((void**)&m->base, total);
        if (e != cudaSuccess) { cudaGetLastError(); m->base = nullptr; ok = 0; }
        else {
            CKC(ctx, cudaMemsetAsync(m->base + payload, 0, sizeof(MgpuCtl), st));
            m->ctl = reinterpret_cast<MgpuCtl*>(m->base + payload);
            if (m->world > 1 && !m->force_nccl && cudaIpcGetMemHandle(&h, m->base) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        }
    }
    m->bytes = bytes;
    m->mode = RXC_MGPU_LOCAL;
    if (m->world > 1) {
        // the handle travels through NCCL (device buffers), then every rank reports whether it could map rank 0's memory
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
        CKC(ctx, cudaMemcpyAsync(m->d_word, &h, sizeof(h), cudaMemcpyHostToDevice, st));
        CKN(ctx, m, m->nccl->Broadcast(m->d_word, m->d_word, sizeof(h), kNcclUint8, 0, m->comm, st));
        CKC(ctx, cudaMemcpyAsync(&h, m->d_word, sizeof(h), cudaMemcpyDeviceToHost, st));
        CKC(ctx, cudaStreamSynchronize(st));
        if (m->force_nccl) ok = 0;
        if (m->rank != 0 && ok) {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { cudaGetLastError(); ok = 0; }
            else { m->base = (uint8_t*)p; m->base_is_ipc = true; m->ctl = reinterpret_cast<MgpuCtl*>(m->base + payload); }
        }
        int32_t all_ok = ok;
        CKC(ctx, cudaMemcpyAsync(m->d_word + 32, &all_ok, 4, cudaMemcpyHostToDevice, st));
        CKN(ctx, m, m->nccl->AllReduce(m->d_word + 32, m->d_word + 32, 1, kNcclInt32, kNcclMin, m->comm, st));
        CKC(ctx, cudaMemcpyAsync(&all_ok, m->d_word + 32, 4, cudaMemcpyDeviceToHost, st));
        CKC(ctx, cudaStreamSynchronize(st));
        if (all_ok) {
            m->mode = m->rank == 0 ? RXC_MGPU_LOCAL : RXC_MGPU_PEER;
        } else {
            // someone has no peer mapping (or rank 0 no memory): everybody stages locally and NCCL moves the regions
            if (m->rank != 0) {
                if (m->base && m->base_is_ipc) cudaIpcCloseMemHandle(m->base);
                m->base = nullptr; m->base_is_ipc = false; m->ctl = nullptr;
                if (cudaMalloc((void**)&m->base, total) != cudaSuccess) { cudaGetLastError(); m->base = nullptr; }
            }
            int32_t have = m->base ? 1 : 0;
            CKC(ctx, cudaMemcpyAsync(m->d_word + 32, &have, 4, cudaMemcpyHostToDevice, st));
            CKN(ctx, m, m->nccl->AllReduce(m->d_word + 32, m->d_word + 32, 1, kNcclInt32, kNcclMin, m->comm, st));
            CKC(ctx, cudaMemcpyAsync(&have, m->d_word + 32, 4, cudaMemcpyDeviceToHost, st));
            CKC(ctx, cudaStreamSynchronize(st));
            if (!have) { free_target(ctx, m); return rxi_fail(ctx, RXC_ERR_OOM, "rxc_mgpu_target: a rank could not allocate the delivery buffer"); }
            m->mode = RXC_MGPU_NCCL;
        }
    } else if (!ok) {
        return rxi_fail(ctx, RXC_ERR_OOM, "rxc_mgpu_target: cudaMalloc of the delivery buffer failed");
    }
    if (rank0_ptr && m->rank == 0) *rank0_ptr = m->base;
    if (mode_out) *mode_out = (m->world > 1 && m->rank == 0 && m->mode != RXC_MGPU_NCCL) ? (uint32_t)RXC_MGPU_LOCAL : m->mode;
    return RXC_OK;
    });
}

int32_t rxc_mgpu_rasterize(rxc_ctx* ctx, const rxc_frame* frames, uint32_t n_frames, uint64_t offset_bytes, uint64_t frame_stride_bytes,
                           uint64_t pitch_bytes) {
    return guarded(ctx, [&]() -> int32_t {
    RxMgpu* m = state(ctx);
    if (!m || !m->base) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_target has not been called");
    if (!frames || n_frames == 0) return rxi_fail(ctx, RXC_ERR_INVALID, "frames are required");
    if ((offset_bytes & 3) || offset_bytes >= m->bytes) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_rasterize: offset outside the delivery buffer");
    // extent of what the call writes: the last row of the last frame
    const rxc_frame& f = frames[0];
    const uint64_t rows = (f.band_y0 | f.band_y1) ? (uint64_t)f.band_y1 - f.band_y0 : f.height;
    const uint64_t cols = (f.band_x0 | f.band_x1) ? (uint64_t)f.band_x1 - f.band_x0 : f.width;
    const uint64_t pitch = pitch_bytes ? pitch_bytes : cols * 4;
    const uint64_t extent = (uint64_t)(n_frames - 1) * frame_stride_bytes + (rows ? rows - 1 : 0) * pitch + cols * 4;
    if (offset_bytes + extent > m->bytes) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_rasterize: the frames do not fit the delivery buffer at that offset");
    CKC(ctx, cudaSetDevice(rxi_device(ctx)));
    if (m->mode == RXC_MGPU_NCCL && m->moved_pending) {   // the staging bytes may still be on their way to rank 0
        CKC(ctx, cudaStreamWaitEvent(rxi_stream(ctx), m->ev_moved, 0));
        m->moved_pending = false;
    }
    return rxi_rasterize_device(ctx, frames, n_frames, m->base + offset_bytes, frame_stride_bytes, pitch_bytes);
    });
}

int32_t rxc_mgpu_deliver(rxc_ctx* ctx, const rxc_mgpu_region* regions, uint32_t n_regions) {
    return guarded(ctx, [&]() -> int32_t {
    RxMgpu* m = state(ctx);
    if (!m || !m->base) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_target has not been called");
    CKC(ctx, cudaSetDevice(rxi_device(ctx)));
    cudaStream_t st = rxi_stream(ctx);
    ++m->seq;
    if (m->world == 1) return RXC_OK;
    if (m->mode != RXC_MGPU_NCCL) {
        if (m->rank != 0) {
            k_mgpu_signal<<<1, 1, 0, st>>>(m->ctl->arrive + m->rank, m->seq);
        } else {
            k_mgpu_wait<<<1, 32, 0, st>>>(m->ctl->arrive, 1u, m->world - 1u, m->seq, &m->ctl->timeouts);
        }
        rxi_count_launch(ctx, 1);
        CKC(ctx, cudaGetLastError());
        return RXC_OK;
    }
    // NCCL mode: the regions of this step, all in one group, on the communication stream
    if (n_regions && !regions) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_deliver: regions are required when peer mappings are unavailable");
    uint64_t pack_need = 0;
    for (uint32_t i = 0; i < n_regions; ++i) {
        const rxc_mgpu_region& g = regions[i];
        if (g.rank >= m->world) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_deliver: region of an unknown rank");
        const uint64_t pitch = g.pitch_bytes ? g.pitch_bytes : g.row_bytes;
        if (g.rows == 0 || g.row_bytes == 0) continue;
        if (pitch < g.row_bytes || g.offset + (uint64_t)(g.rows - 1) * pitch + g.row_bytes > m->bytes) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_deliver: region outside the delivery buffer");
        if (pitch != g.row_bytes && g.rank != 0 && (m->rank == 0 || m->rank == g.rank)) pack_need += ((uint64_t)g.rows * g.row_bytes + 255) & ~(uint64_t)255;
    }
    if (pack_need > m->pack_cap) {
        CKC(ctx, cudaStreamSynchronize(m->comm_stream));
        if (m->pack) cudaFree(m->pack);
        m->pack = nullptr; m->pack_cap = 0;
        if (cudaMalloc((void**)&m->pack, pack_need) != cudaSuccess) { cudaGetLastError(); return rxi_fail(ctx, RXC_ERR_OOM, "rxc_mgpu_deliver: staging for pitched regions"); }
        m->pack_cap = pack_need;
    }
    CKC(ctx, cudaEventRecord(m->ev_rendered, st));
    CKC(ctx, cudaStreamWaitEvent(m->comm_stream, m->ev_rendered, 0));
    uint64_t po = 0;
    std::vector<std::pair<const rxc_mgpu_region*, uint64_t>> unpack;
    for (uint32_t i = 0; i < n_regions; ++i) {   // senders pack their pitched regions first
        const rxc_mgpu_region& g = regions[i];
        const uint64_t pitch = g.pitch_bytes ? g.pitch_bytes : g.row_bytes;
        if (g.rows == 0 || g.row_bytes == 0 || g.rank == 0 || pitch == g.row_bytes || !(m->rank == 0 || m->rank == g.rank)) continue;
        if (m->rank == g.rank) CKC(ctx, cudaMemcpy2DAsync(m->pack + po, g.row_bytes, m->base + g.offset, pitch, g.row_bytes, g.rows, cudaMemcpyDeviceToDevice, m->comm_stream));
        else unpack.push_back({&g, po});
        po += ((uint64_t)g.rows * g.row_bytes + 255) & ~(uint64_t)255;
    }
    CKN(ctx, m, m->nccl->GroupStart());
    po = 0;
    for (uint32_t i = 0; i < n_regions; ++i) {
        const rxc_mgpu_region& g = regions[i];
        const uint64_t pitch = g.pitch_bytes ? g.pitch_bytes : g.row_bytes;
        if (g.rows == 0 || g.row_bytes == 0 || g.rank == 0 || !(m->rank == 0 || m->rank == g.rank)) continue;
        const uint64_t n = (uint64_t)g.rows * g.row_bytes;
        uint8_t* p = m->base + g.offset;
        if (pitch != g.row_bytes) { p = m->pack + po; po += (n + 255) & ~(uint64_t)255; }
        if (m->rank == 0) CKN(ctx, m, m->nccl->Recv(p, n, kNcclUint8, (int)g.rank, m->comm, m->comm_stream));
        else CKN(ctx, m, m->nccl->Send(p, n, kNcclUint8, 0, m->comm, m->comm_stream));
    }
    CKN(ctx, m, m->nccl->GroupEnd());
    for (auto& u : unpack) {
        const rxc_mgpu_region& g = *u.first;
        CKC(ctx, cudaMemcpy2DAsync(m->base + g.offset, g.pitch_bytes, m->pack + u.second, g.row_bytes, g.row_bytes, g.rows, cudaMemcpyDeviceToDevice, m->comm_stream));
    }
    CKC(ctx, cudaEventRecord(m->ev_moved, m->comm_stream));
    if (m->rank == 0) CKC(ctx, cudaStreamWaitEvent(st, m->ev_moved, 0));   // what follows on rank 0's stream sees the regions
    else m->moved_pending = true;                                          // the next render into the staging waits for the sends
    return RXC_OK;
    });
}

int32_t rxc_mgpu_release(rxc_ctx* ctx) {
    return guarded(ctx, [&]() -> int32_t {
    RxMgpu* m = state(ctx);
    if (!m || !m->base) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_target has not been called");
    CKC(ctx, cudaSetDevice(rxi_device(ctx)));
    ++m->rseq;
    if (m->world == 1 || m->mode == RXC_MGPU_NCCL) return RXC_OK;   // NCCL mode: rank 0 posts its receives when it is ready for them
    cudaStream_t st = rxi_stream(ctx);
    if (m->rank == 0) k_mgpu_signal<<<1, 1, 0, st>>>(&m->ctl->release, m->rseq);
    else k_mgpu_wait<<<1, 32, 0, st>>>(&m->ctl->release, 0u, 1u, m->rseq, &m->ctl->timeouts);
    rxi_count_launch(ctx, 1);
    CKC(ctx, cudaGetLastError());
    return RXC_OK;
    });
}

int32_t rxc_mgpu_status(rxc_ctx* ctx, uint32_t* mode, uint32_t* deliveries, uint32_t* timeouts) {
    return guarded(ctx, [&]() -> int32_t {
    RxMgpu* m = state(ctx);
    if (!m) return rxi_fail(ctx, RXC_ERR_INVALID, "rxc_mgpu_init has not been called");
    CKC(ctx, cudaSetDevice(rxi_device(ctx)));
    if (mode) *mode = m->mode;
    if (deliveries) *deliveries = m->seq;
    if (timeouts) {
        *timeouts = 0;
        if (m->ctl && m->rank == 0) {
            CKC(ctx, cudaStreamSynchronize(rxi_stream(ctx)));
            CKC(ctx, cudaMemcpy(timeouts, &m->ctl->timeouts, 4, cudaMemcpyDeviceToHost));
        }
    }
    return RXC_OK;
    });
}

}  // extern "C"
