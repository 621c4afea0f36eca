// comm.cu -- z-slab decomposition plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
//
// The reference has no multi-device code at all; this file is what the slab
// decomposition adds around the single-GPU kernels:
//   * halo exchange of the z ghost planes (replaces the periodic copy of apply_bc_u!/p!/temp!
//     along the decomposed direction, src/boundary_conditions.jl:273-285,303-315),
//   * the all-to-all transposes of the distributed FFT Poisson solve (z-slabs <-> ky-slabs),
//   * min/max reductions for propose_timestep / maximum(abs, u).
// NCCL is loaded lazily with dlopen so that a single-GPU run never depends on it; when torch
// is already in the process its bundled libnccl.so.2 is the one that gets bound.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "internal.h"

namespace insb {

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error(std::string("cannot load libnccl.so.2 (needed for world > 1): ") + dlerror());
        return 1;
    }
#define LOAD(field, sym)                                               \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, sym)); \
    if (!g_nccl.field) {                                               \
        set_error(std::string("libnccl lacks ") + sym);                \
        return 1;                                                      \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(GetErrorString, "ncclGetErrorString")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(AllReduce, "ncclAllReduce")
#undef LOAD
    g_nccl.handle = h;
    return 0;
}

#define INSB_NCCL(call)                                                                     \
    do {                                                                                    \
        ncclResult_t r__ = (call);                                                          \
        if (r__ != ncclSuccess) {                                                           \
            set_error(std::string(#call) + ": " + g_nccl.GetErrorString(r__) + " at " +     \
                      __FILE__ + ":" + std::to_string(__LINE__));                           \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

}  // namespace

// How long a rank waits for a peer's flag before it gives up with a sticky launch error (__trap): 20 s, or
// INSB200_PEER_TIMEOUT_MS (tests/mgpu_worker.py --case timeout uses a short one).
static unsigned long long peer_timeout_ns() {
    static unsigned long long ns = 0;
    if (!ns) {
        const char *e = getenv("INSB200_PEER_TIMEOUT_MS");
        const long long ms = e ? atoll(e) : 20000;
        ns = (unsigned long long)(ms > 0 ? ms : 20000) * 1000000ULL;
    }
    return ns;
}

struct Comm {
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    int refs = 1;
    std::string key;
    // direct peer-store halo exchange (comm_enable_p2p_halo): every rank owns a window
    // [flag words | staging slots] in peer-mapped memory
    bool halo_p2p = false;
    char *halo_win[8] = {};
    size_t halo_slot = 0;  // bytes per staging slot
    unsigned halo_epoch = 0;
};

// window layout: unsigned words 0..23 belong to the transposes (poisson.cu), word 16 is the
// time-out marker; halo flags at HALO_FLAG0 + parity*2 + dir, block counter at HALO_COUNTER;
// staging slots (parity, dir) start at byte HALO_STAGING
constexpr int HALO_FLAG0 = 32, HALO_COUNTER = 48, HALO_TIMEOUT = 16;
constexpr size_t HALO_STAGING = 4096;
size_t comm_halo_window_bytes(size_t slot_bytes) { return HALO_STAGING + 4 * ((slot_bytes + 255) / 256 * 256); }

// One communicator per unique id and process: several plans of the same setup (base plan,
// psolver plans) share it, since an id can be passed to ncclCommInitRank only once.
static std::map<std::string, Comm *> g_comms;
static std::mutex g_comms_mu;

int comm_unique_id(void *id128) {
    INSB_TRY(load_nccl());
    ncclUniqueId id;
    INSB_NCCL(g_nccl.GetUniqueId(&id));
    std::memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return 0;
}

int comm_create(int rank, int world, const void *id128, Comm **out) {
    *out = nullptr;
    if (world < 2 || rank < 0 || rank >= world) {
        set_error("comm_create: bad (rank, world)");
        return 1;
    }
    if (!id128) {
        set_error("world > 1 needs setup.nccl_id (insb200_comm_unique_id on rank 0, shared with all ranks)");
        return 1;
    }
    INSB_TRY(load_nccl());
    ncclUniqueId id;
    std::memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    std::lock_guard<std::mutex> lock(g_comms_mu);
    std::string key((const char *)id128, NCCL_UNIQUE_ID_BYTES);
    auto it = g_comms.find(key);
    if (it != g_comms.end()) {
        if (it->second->rank != rank || it->second->world != world) {
            set_error("comm_create: nccl_id already in use with a different (rank, world)");
            return 1;
        }
        it->second->refs++;
        *out = it->second;
        return 0;
    }
    Comm *c = new Comm();
    c->key = key;
    c->rank = rank;
    c->world = world;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
        delete c;
        return 1;
    }
    g_comms[key] = c;
    *out = c;
    return 0;
}

void comm_destroy(Comm *c) {
    if (!c) return;
    std::lock_guard<std::mutex> lock(g_comms_mu);
    if (--c->refs > 0) return;
    g_comms.erase(c->key);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
}

int comm_rank(const Comm *c) { return c ? c->rank : 0; }
int comm_world(const Comm *c) { return c ? c->world : 1; }

// Ghost planes of the decomposed (slowest) direction.  A field component is `nplanes`
// contiguous planes of `plane` elements; plane 0 and plane nplanes-1 are the ghosts.
//   lower ghost  <- top interior plane of the rank below      (u[1]   = u[N-1] in the reference)
//   upper ghost  <- first interior plane of the rank above    (u[N]   = u[2])
// Posting order (send up, send down, recv from below, recv from above) keeps the pairs
// matched when world == 2 and both neighbours are the same peer.
// ---- halo exchange by direct peer stores --------------------------------------------------
// push: the top interior plane of every component goes into the upper neighbour's "from below"
// slot, the first interior plane into the lower neighbour's "from above" slot; the last block
// to finish raises the neighbours' flags (stores -> __threadfence_system -> counter -> flags).
// pull: waits for both flags of this epoch, then copies the slots into the ghost planes.
// Slots and flags are double-buffered by the parity of the exchange count: a rank can be at
// most one exchange ahead of its neighbours (it needs their data of exchange e before it can
// start e + 1), so the slot of exchange e + 2 is free again when it is written.
template <typename V>
__global__ void __launch_bounds__(256) halo_push_kernel(const char *src_up, const char *src_down, size_t chunk_bytes,
                                                        size_t cs_bytes, char *up_slot, char *down_slot,
                                                        unsigned *up_flag, unsigned *down_flag, unsigned epoch,
                                                        unsigned *counter) {
    const int comp = blockIdx.y, dir = blockIdx.z;  // dir 0: send up, 1: send down
    const size_t nv = chunk_bytes / sizeof(V);
    const char *s = dir == 0 ? src_up : src_down;
    if (s) {
        const V *src = reinterpret_cast<const V *>(s + comp * cs_bytes);
        V *dst = reinterpret_cast<V *>((dir == 0 ? up_slot : down_slot) + comp * chunk_bytes);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x)
            dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(counter, 1u) == total - 1) {
            *counter = 0;
            __threadfence_system();
            *(volatile unsigned *)up_flag = epoch;
            *(volatile unsigned *)down_flag = epoch;
            __threadfence_system();
        }
    }
}

template <typename V>
__global__ void __launch_bounds__(256) halo_pull_kernel(char *dst_below, char *dst_above, size_t chunk_bytes,
                                                        size_t cs_bytes, const char *below_slot, const char *above_slot,
                                                        const unsigned *below_flag, const unsigned *above_flag,
                                                        unsigned epoch, unsigned *timeout_flag,
                                                        unsigned long long timeout_ns) {
    const int comp = blockIdx.y, dir = blockIdx.z;  // dir 0: data from below -> lower ghost, 1: from above -> upper ghost
    if (threadIdx.x == 0) {
        const volatile unsigned *f = dir == 0 ? below_flag : above_flag;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(*f - epoch) < 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {  // (20 s by default) a peer died; fail loudly instead of hanging or
                *timeout_flag = 1;           // continuing with stale ghost planes
                __threadfence_system();
                __trap();
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    char *d = dir == 0 ? dst_below : dst_above;
    if (!d) return;
    const size_t nv = chunk_bytes / sizeof(V);
    const V *src = reinterpret_cast<const V *>((dir == 0 ? below_slot : above_slot) + comp * chunk_bytes);
    V *dst = reinterpret_cast<V *>(d + comp * cs_bytes);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __ldcv(src + i);  // written by a peer: do not serve from a stale cache line
}

// chunk `send_up` (+ c * src_cs) of every component c goes to the rank above, which receives it in `recv_below`
// (+ c * dst_cs); `send_down` / `recv_above` likewise for the other direction.  Null pointers skip a direction
// (all ranks alike); the flags of both directions are raised and awaited regardless.
template <typename V>
static int exchange_p2p_t(Comm *c, const char *send_up, const char *send_down, char *recv_below, char *recv_above,
                          size_t chunk, int ncomp, size_t src_cs, size_t dst_cs, cudaStream_t st) {
    const int up = (c->rank + 1) % c->world, down = (c->rank + c->world - 1) % c->world;
    const unsigned e = ++c->halo_epoch;
    const int q = e & 1;
    auto slot = [&](int r, int dir) { return c->halo_win[r] + HALO_STAGING + (size_t)(q * 2 + dir) * c->halo_slot; };
    auto flag = [&](int r, int dir) { return reinterpret_cast<unsigned *>(c->halo_win[r]) + HALO_FLAG0 + q * 2 + dir; };
    unsigned *mine = reinterpret_cast<unsigned *>(c->halo_win[c->rank]);
    const size_t nv = chunk / sizeof(V);
    unsigned nb = (unsigned)((nv + 255) / 256);
    if (nb > 64) nb = 64;
    dim3 gr(nb, ncomp, 2);
    halo_push_kernel<V><<<gr, 256, 0, st>>>(send_up, send_down, chunk, src_cs, slot(up, 0), slot(down, 1), flag(up, 0),
                                            flag(down, 1), e, mine + HALO_COUNTER);
    INSB_CUDA(cudaGetLastError());
    halo_pull_kernel<V><<<gr, 256, 0, st>>>(recv_below, recv_above, chunk, dst_cs, slot(c->rank, 0), slot(c->rank, 1),
                                            flag(c->rank, 0), flag(c->rank, 1), e, mine + HALO_TIMEOUT, peer_timeout_ns());
    INSB_CUDA(cudaGetLastError());
    return 0;
}

int comm_enable_p2p_halo(Comm *c, void *const *windows, size_t slot_bytes) {
    if (!c || c->halo_p2p) return 0;  // first registration wins (all ranks register in the same order)
    for (int r = 0; r < c->world; r++) c->halo_win[r] = (char *)windows[r];
    c->halo_slot = (slot_bytes + 255) / 256 * 256;
    c->halo_p2p = getenv("INSB200_NCCL_HALO") == nullptr;
    return 0;
}

void comm_disable_p2p_halo(Comm *c, const void *my_window) {
    if (c && c->halo_p2p && c->halo_win[c->rank] == (const char *)my_window) c->halo_p2p = false;
}

int comm_halo_launches(const Comm *c) { return c && c->halo_p2p ? 2 : 0; }

int comm_exchange(Comm *c, const void *send_up, const void *send_down, void *recv_below, void *recv_above,
                  size_t chunk_bytes, int ncomp, size_t src_cs_bytes, size_t dst_cs_bytes, cudaStream_t st) {
    if (c->halo_p2p && (size_t)ncomp * chunk_bytes <= c->halo_slot) {
        const uintptr_t all = (uintptr_t)send_up | (uintptr_t)send_down | (uintptr_t)recv_below | (uintptr_t)recv_above |
                              chunk_bytes | src_cs_bytes | dst_cs_bytes;
        if (all % 16 == 0)
            return exchange_p2p_t<uint4>(c, (const char *)send_up, (const char *)send_down, (char *)recv_below,
                                         (char *)recv_above, chunk_bytes, ncomp, src_cs_bytes, dst_cs_bytes, st);
        return exchange_p2p_t<unsigned>(c, (const char *)send_up, (const char *)send_down, (char *)recv_below,
                                        (char *)recv_above, chunk_bytes, ncomp, src_cs_bytes, dst_cs_bytes, st);
    }
    // Posting order (send up, send down, recv from below, recv from above) keeps the pairs matched when
    // world == 2 and both neighbours are the same peer.
    const int up = (c->rank + 1) % c->world, down = (c->rank + c->world - 1) % c->world;
    INSB_NCCL(g_nccl.GroupStart());
    for (int k = 0; k < ncomp; k++) {
        const size_t so = (size_t)k * src_cs_bytes, dof = (size_t)k * dst_cs_bytes;
        if (send_up) INSB_NCCL(g_nccl.Send((const char *)send_up + so, chunk_bytes, ncclChar, up, c->comm, st));
        if (send_down) INSB_NCCL(g_nccl.Send((const char *)send_down + so, chunk_bytes, ncclChar, down, c->comm, st));
        if (recv_below) INSB_NCCL(g_nccl.Recv((char *)recv_below + dof, chunk_bytes, ncclChar, down, c->comm, st));
        if (recv_above) INSB_NCCL(g_nccl.Recv((char *)recv_above + dof, chunk_bytes, ncclChar, up, c->comm, st));
    }
    INSB_NCCL(g_nccl.GroupEnd());
    return 0;
}

int comm_exchange_launches(const Comm *c, size_t chunk_bytes, int ncomp) {
    return c && c->halo_p2p && (size_t)ncomp * chunk_bytes <= c->halo_slot ? 2 : 0;
}

// Ghost planes of a field whose component is `nplanes` contiguous planes of `plane` elements with `ng` ghost
// planes on either side: the top ng interior planes go up, the first ng interior planes go down.
// ring == false: the decomposed direction is not periodic -- nothing travels between the last and the first rank
// (their outer ghost planes belong to the physical boundary conditions and stay untouched).
int comm_halo_planes(Comm *c, void *field, size_t esz, long long plane, int nplanes, int ng, int ncomp, long long cs,
                     cudaStream_t st, bool ring) {
    char *f = (char *)field;
    const size_t pb = (size_t)plane * esz, csb = (size_t)cs * esz;
    const bool first = c->rank == 0, last = c->rank == c->world - 1;
    const bool up = ring || !last, down = ring || !first;
    return comm_exchange(c, up ? f + (size_t)(nplanes - 2 * ng) * pb : nullptr, down ? f + (size_t)ng * pb : nullptr,
                         down ? f : nullptr, up ? f + (size_t)(nplanes - ng) * pb : nullptr, (size_t)ng * pb, ncomp, csb,
                         csb, st);
}

int comm_halo_z(Comm *c, void *field, size_t esz, long long plane, int nplanes, int ncomp, long long cs,
                cudaStream_t st, bool ring) {
    return comm_halo_planes(c, field, esz, plane, nplanes, 1, ncomp, cs, st, ring);
}

// All-to-all of `world` equal blocks of `block_bytes`: block r of `send` goes to rank r and
// lands as block `rank` of its `recv`.
int comm_all_to_all(Comm *c, const void *send, void *recv, size_t block_bytes, cudaStream_t st) {
    const char *s = (const char *)send;
    char *r = (char *)recv;
    INSB_CUDA(cudaMemcpyAsync(r + (size_t)c->rank * block_bytes, s + (size_t)c->rank * block_bytes, block_bytes,
                              cudaMemcpyDeviceToDevice, st));
    INSB_NCCL(g_nccl.GroupStart());
    for (int d = 1; d < c->world; d++) {
        int to = (c->rank + d) % c->world, from = (c->rank + c->world - d) % c->world;
        INSB_NCCL(g_nccl.Send(s + (size_t)to * block_bytes, block_bytes, ncclChar, to, c->comm, st));
        INSB_NCCL(g_nccl.Recv(r + (size_t)from * block_bytes, block_bytes, ncclChar, from, c->comm, st));
    }
    INSB_NCCL(g_nccl.GroupEnd());
    return 0;
}

// scratch[0]: non-negative double reduced with min or max; scratch[1]: NaN flag (max)
int comm_reduce_scratch(Comm *c, double *scratch, bool is_min, cudaStream_t st) {
    INSB_NCCL(g_nccl.GroupStart());
    INSB_NCCL(g_nccl.AllReduce(scratch, scratch, 1, ncclDouble, is_min ? ncclMin : ncclMax, c->comm, st));
    INSB_NCCL(g_nccl.AllReduce(scratch + 1, scratch + 1, 1, ncclUint64, ncclMax, c->comm, st));
    INSB_NCCL(g_nccl.GroupEnd());
    return 0;
}

int comm_sum_double(Comm *c, double *value, cudaStream_t st) {
    INSB_NCCL(g_nccl.AllReduce(value, value, 1, ncclDouble, ncclSum, c->comm, st));
    return 0;
}

int comm_sum_doubles(Comm *c, double *values, int n, cudaStream_t st) {
    INSB_NCCL(g_nccl.AllReduce(values, values, (size_t)n, ncclDouble, ncclSum, c->comm, st));
    return 0;
}

// ---- peer flags ---------------------------------------------------------------------------
struct PeerPtrs {
    unsigned *p[8];
};

__global__ void signal_peers_kernel(PeerPtrs flags, int row, int rank, int world, unsigned epoch) {
    __threadfence_system();
    int s = threadIdx.x;
    if (s < world) {
        volatile unsigned *f = flags.p[s] + row * 8 + rank;
        *f = epoch;
    }
    __threadfence_system();
}

__global__ void wait_flags_kernel(unsigned *flags, int row, int world, unsigned epoch, unsigned *timeout_flag,
                                  unsigned long long timeout_ns) {
    int s = threadIdx.x;
    if (s < world) {
        volatile unsigned *f = flags + row * 8 + s;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(*f - epoch) < 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {  // (20 s by default) a peer died; fail loudly (sticky launch error)
                *timeout_flag = 1;           // instead of hanging or transforming stale data
                __threadfence_system();
                __trap();
            }
        }
    }
    __threadfence_system();
}

int comm_signal_peers(unsigned *const *peer_flags, int row, int rank, int world, unsigned epoch, cudaStream_t st) {
    PeerPtrs pp{};
    for (int s = 0; s < world; s++) pp.p[s] = peer_flags[s];
    signal_peers_kernel<<<1, 32, 0, st>>>(pp, row, rank, world, epoch);
    INSB_CUDA(cudaGetLastError());
    return 0;
}

int comm_wait_flags(unsigned *my_flags, int row, int world, unsigned epoch, unsigned *timeout_flag, cudaStream_t st) {
    wait_flags_kernel<<<1, 32, 0, st>>>(my_flags, row, world, epoch, timeout_flag, peer_timeout_ns());
    INSB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace insb
