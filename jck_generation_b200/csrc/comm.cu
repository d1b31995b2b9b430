// Peer-memory communicator for the small, latency-bound exchanges of data-parallel training: the SyncBN
// statistics [sum x, sum x^2] / [sum g, sum g*xhat] (2C fp32 per layer per pass, 40+ exchanges per step).
//
// One process per GPU.  Every rank cudaMalloc's one mailbox region, exports it with CUDA IPC and maps its
// peers' regions, so a kernel can store straight into a peer's HBM over NVLink / NVSwitch.  An all-reduce of
// n <= JCK_COMM_MAX_N floats is then ONE single-CTA kernel per rank ("one-shot", push model):
//     for every peer p:  mailbox_p[slot][my_rank][i] = (value_i, seq)      one 8-byte store: data + flag together
//     for every rank r:  spin on my own mailbox[slot][r][i] until its flag == seq;  sum in rank order
// -- one NVLink one-way latency, no host involvement, no separate barrier, bit-identical sums on all ranks
// (fixed summation order).  `seq` is a device-side counter the kernel advances itself, so the launch sequence
// is CUDA-graph capturable.  A slot is reused after kSlots calls; a rank can only be kSlots calls ahead of a
// peer if that peer has delivered its values for the calls in between, i.e. has finished reading the slot.
// Every spin is bounded (JCK_COMM_TIMEOUT_S seconds, default 120 -- generous, like NCCL's watchdog: legitimate rank
// skew can be seconds when rank 0 alone writes a checkpoint or runs an evaluation while its peers already wait in the
// next step's exchange).  A timeout does NOT trap (that would destroy the CUDA context of a job that may be recoverable):
// the kernel prints a diagnostic, raises a sticky error flag the host reads with jck_comm_error(), and poisons its result
// with NaN so that nothing downstream can mistake it for a sum.
//
// Programmatic dependent launch and spin-waits on a REMOTE peer: a kernel that has executed launch_dependents lets its
// dependents become resident early; they then sit in griddepcontrol.wait holding registers / thread slots on every SM
// while this kernel waits for the peer.  With ONE stream of exchanges that is harmless (the peer's counterpart only has
// ordinary kernels in front of it).  With TWO streams issuing exchanges (the penalty sweep runs on its own stream with its
// own communicator, parallel.AuxComm) it deadlocks across ranks: rank 0 spins in a main-stream exchange whose dependents
// fill the GPU, so its auxiliary exchange cannot get an SM, while rank 1 spins in the auxiliary exchange whose dependents
// starve its main one (seen at 2 GPUs with 1024-thread exchange kernels).  So by default a communicator NEVER triggers
// early: its dependents start only when the exchange has completed (jck_comm_configure can turn early launch on for ONE
// communicator of a process; measured, it does not pay).  The kernels are 512 threads (half an SM's registers) for the same reason.
#include <stdlib.h>
#include "common.cuh"

namespace jck {
namespace {

constexpr int kSlots = 4;
constexpr int kMaxWorld = 8;
constexpr int kMaxN = JCK_COMM_MAX_N;

struct CommDev {
    unsigned long long* peer[kMaxWorld];   // mailbox base of every rank (own included), as mapped in THIS process
    unsigned int* seq;                     // local call counter
    unsigned int* err;                     // sticky error flag (device memory, read back by jck_comm_error)
    long long timeout_cycles;              // spin bound
    int rank, world;
    int early_dependents;                  // execute griddepcontrol.launch_dependents on entry (see above)
};

__device__ __forceinline__ void comm_entry(const CommDev& c) {
    if (c.early_dependents) pdl_trigger();
    pdl_wait();
}

struct Comm {
    CommDev dev;
    void* local;                           // cudaMalloc'd mailbox of this rank
    void* opened[kMaxWorld];               // cudaIpcOpenMemHandle results (peers)
    int device;
};

__host__ __device__ inline size_t mailbox_words() { return (size_t)kSlots * kMaxWorld * kMaxN; }

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// In-place sum over ranks of data[0..n), executed by ONE thread block; every thread must call it.
__device__ void allreduce_block(const CommDev& c, float* __restrict__ data, int n) {
    const unsigned int seq = *c.seq + 1u;
    const int slot = (int)(seq % kSlots);
    const size_t mine = ((size_t)slot * kMaxWorld + c.rank) * kMaxN;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long word = ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(data[i]);
        for (int p = 0; p < c.world; ++p) st_sys_u64(c.peer[p] + mine + i, word);
    }
    const unsigned long long* box = c.peer[c.rank] + (size_t)slot * kMaxWorld * kMaxN;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < c.world; ++r) {
            const unsigned long long* w = box + (size_t)r * kMaxN + i;
            unsigned long long v = ld_sys_u64(w);
            if ((unsigned int)(v >> 32) != seq) {
                const long long t0 = clock64();
                do {
                    if (clock64() - t0 > c.timeout_cycles) {   // a peer never arrived
                        if (atomicExch(c.err, 1u) == 0u)
                            printf("jck: comm all-reduce timed out (rank %d waits for rank %d, call %u)\n", c.rank, r, seq);
                        v = ((unsigned long long)seq << 32) | 0x7FC00000ull;   // NaN: poison, do not pretend
                        break;
                    }
                    v = ld_sys_u64(w);
                } while ((unsigned int)(v >> 32) != seq);
            }
            s += __uint_as_float((unsigned int)v);
        }
        data[i] = s;
    }
    __syncthreads();                       // every thread has read *c.seq and finished its part
    if (threadIdx.x == 0) *c.seq = seq;
}

constexpr int kCommThreads = 512;

__global__ void __launch_bounds__(kCommThreads) comm_allreduce_kernel(const CommDev c, float* __restrict__ data, int n) {
    comm_entry(c);
    allreduce_block(c, data, n);
}

// SyncBN forward: exchange [groups][2C] statistics, then what jck_bn_finalize does (same arithmetic).
__global__ void __launch_bounds__(kCommThreads)
bn_finalize_sync_kernel(const CommDev c, float* __restrict__ stats, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                        long long* __restrict__ nbt, float* __restrict__ scale_shift, float* __restrict__ mean_rstd, int C,
                        int groups, float count, float eps, float momentum) {
    comm_entry(c);
    allreduce_block(c, stats, groups * 2 * C);
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
        float rm = running_mean ? running_mean[ch] : 0.f, rv = running_var ? running_var[ch] : 0.f;
        for (int g = 0; g < groups; ++g) {
            const double s1 = stats[(size_t)g * 2 * C + ch], s2 = stats[(size_t)g * 2 * C + C + ch];
            const double mean = s1 / count;
            double var = s2 / count - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + (double)eps));
            const float sc = gamma[ch] * rstd;
            scale_shift[(size_t)g * 2 * C + ch] = sc;
            scale_shift[(size_t)g * 2 * C + C + ch] = beta[ch] - (float)mean * sc;
            mean_rstd[(size_t)g * 2 * C + ch] = (float)mean;
            mean_rstd[(size_t)g * 2 * C + C + ch] = rstd;
            const float unbiased = (float)(count > 1.f ? var * (double)count / ((double)count - 1.0) : var);
            rm = (1.f - momentum) * rm + momentum * (float)mean;
            rv = (1.f - momentum) * rv + momentum * unbiased;
        }
        if (running_mean) running_mean[ch] = rm;
        if (running_var) running_var[ch] = rv;
    }
    if (nbt && threadIdx.x == 0) *nbt += groups;
}

// SyncBN backward: this rank's share of dgamma / dbeta from its LOCAL sums, then exchange the sums.
__global__ void __launch_bounds__(kCommThreads)
bn_bwd_sums_sync_kernel(const CommDev c, float* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta,
                        int C, int groups, int accumulate) {
    comm_entry(c);
    if (dgamma != nullptr) {
        for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
            float sb = 0.f, sg = 0.f;
            for (int g = 0; g < groups; ++g) { sb += sums[(size_t)g * 2 * C + ch]; sg += sums[(size_t)g * 2 * C + C + ch]; }
            dgamma[ch] = accumulate ? dgamma[ch] + sg : sg;
            dbeta[ch] = accumulate ? dbeta[ch] + sb : sb;
        }
        __syncthreads();
    }
    allreduce_block(c, sums, groups * 2 * C);
}

}  // namespace
}  // namespace jck

using namespace jck;

extern "C" int jck_comm_create(int rank, int world, void** comm_out, void* ipc_handle_out) {
    JCK_REQUIRE(comm_out && ipc_handle_out && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world,
                "comm_create: bad argument (world <= %d)", kMaxWorld);
    static_assert(sizeof(cudaIpcMemHandle_t) == JCK_COMM_HANDLE_BYTES, "IPC handle size");
    Comm* cm = new Comm();
    cm->dev.rank = rank;
    cm->dev.world = world;
    cm->dev.early_dependents = 0;          // measured at 2 GPUs: early launch buys nothing here (2.82 vs 2.75 ms/step) and is unsafe with two streams
    cudaError_t e = cudaGetDevice(&cm->device);
    const size_t bytes = mailbox_words() * sizeof(unsigned long long) + 256;
    {
        const char* t = getenv("JCK_COMM_TIMEOUT_S");
        double secs = t ? atof(t) : 120.0;
        if (!(secs > 0.0)) secs = 120.0;
        cm->dev.timeout_cycles = (long long)(secs * 2.0e9);            // clock64 ticks at ~2 GHz
    }
    if (e == cudaSuccess) e = cudaMalloc(&cm->local, bytes);
    if (e == cudaSuccess) e = cudaMemset(cm->local, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, cm->local);
    if (e != cudaSuccess) {
        delete cm;
        return set_error(JCK_E_CUDA, "comm_create: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle_out, &h, sizeof(h));
    cm->dev.peer[rank] = reinterpret_cast<unsigned long long*>(cm->local);
    cm->dev.seq = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(cm->local) + mailbox_words() * sizeof(unsigned long long));
    cm->dev.err = cm->dev.seq + 16;
    *comm_out = cm;
    return JCK_OK;
}

extern "C" int jck_comm_connect(void* comm, const void* all_handles) {
    JCK_REQUIRE(comm && all_handles, "comm_connect: bad argument");
    Comm* cm = static_cast<Comm*>(comm);
    for (int r = 0; r < cm->dev.world; ++r) {
        if (r == cm->dev.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(all_handles) + (size_t)r * JCK_COMM_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return set_error(JCK_E_CUDA, "comm_connect: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        cm->opened[r] = p;
        cm->dev.peer[r] = reinterpret_cast<unsigned long long*>(p);
    }
    return JCK_OK;
}

extern "C" int jck_comm_configure(void* comm, int early_dependents) {
    JCK_REQUIRE(comm, "comm_configure: bad argument");
    static_cast<Comm*>(comm)->dev.early_dependents = early_dependents ? 1 : 0;
    return JCK_OK;
}

extern "C" int jck_comm_error(void* comm, int* flag_out) {
    JCK_REQUIRE(comm && flag_out, "comm_error: bad argument");
    Comm* cm = static_cast<Comm*>(comm);
    unsigned int f = 0;
    cudaError_t e = cudaMemcpy(&f, cm->dev.err, sizeof(f), cudaMemcpyDeviceToHost);      // synchronises with the device
    if (e != cudaSuccess) return set_error(JCK_E_CUDA, "comm_error: %s", cudaGetErrorString(e));
    *flag_out = (int)f;
    return JCK_OK;
}

extern "C" int jck_comm_destroy(void* comm) {
    if (!comm) return JCK_OK;
    Comm* cm = static_cast<Comm*>(comm);
    for (int r = 0; r < cm->dev.world; ++r)
        if (r != cm->dev.rank && cm->opened[r]) cudaIpcCloseMemHandle(cm->opened[r]);
    cudaFree(cm->local);
    delete cm;
    return JCK_OK;
}

extern "C" int jck_comm_allreduce_small(void* comm, float* data, int n, void* stream) {
    JCK_REQUIRE(comm && data && n > 0 && n <= kMaxN, "comm_allreduce_small: n=%d (max %d)", n, kMaxN);
    Comm* cm = static_cast<Comm*>(comm);
    launch_pdl(comm_allreduce_kernel, dim3(1), dim3(kCommThreads), 0, as_stream(stream), cm->dev, data, n);
    JCK_LAUNCH_CHECK("comm_allreduce_small");
    return JCK_OK;
}

extern "C" int jck_bn_finalize_sync(void* comm, float* stats, const float* gamma, const float* beta, float* running_mean,
                                    float* running_var, long long* num_batches_tracked, float* scale_shift,
                                    float* mean_rstd, int C, int groups, float count, float eps, float momentum, void* stream) {
    JCK_REQUIRE(comm && stats && gamma && beta && scale_shift && mean_rstd && C > 0 && groups > 0 && count > 0 &&
                groups * 2 * C <= kMaxN, "bn_finalize_sync: bad argument");
    Comm* cm = static_cast<Comm*>(comm);
    launch_pdl(bn_finalize_sync_kernel, dim3(1), dim3(kCommThreads), 0, as_stream(stream), cm->dev, stats, gamma, beta, running_mean, running_var,
                                                             num_batches_tracked, scale_shift, mean_rstd, C, groups, count,
                                                             eps, momentum);
    JCK_LAUNCH_CHECK("bn_finalize_sync");
    return JCK_OK;
}

extern "C" int jck_bn_bwd_sums_sync(void* comm, float* sums, float* dgamma, float* dbeta, int C, int groups, int accumulate,
                                    void* stream) {
    JCK_REQUIRE(comm && sums && C > 0 && groups > 0 && groups * 2 * C <= kMaxN && ((dgamma == nullptr) == (dbeta == nullptr)),
                "bn_bwd_sums_sync: bad argument");
    Comm* cm = static_cast<Comm*>(comm);
    launch_pdl(bn_bwd_sums_sync_kernel, dim3(1), dim3(kCommThreads), 0, as_stream(stream), cm->dev, sums, dgamma, dbeta, C, groups, accumulate);
    JCK_LAUNCH_CHECK("bn_bwd_sums_sync");
    return JCK_OK;
}
