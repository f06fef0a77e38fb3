// flgpu_exchange.cuh -- the rank exchange over NVSwitch peer memory as device code (libflgpu's exchange_kernel and
// device-resident line search; the device-resident search that include/flgpu_objective.cuh generates for user functors).
//
// Every rank owns one IPC-shared allocation with two mailboxes; a mailbox has, per sequence parity, one slot of
// kMailWidth doubles and one flag word per rank.  An exchange: store this rank's values straight into slot [me] of every
// peer's mailbox (plain st.global on IPC-mapped pointers), publish the sequence number with st.release.sys, spin with
// ld.acquire.sys until the G flags of the own mailbox carry it, combine the G slots by the aligned rank tree
// (flgpu_reduce_geom.h) -- identical bits on all ranks.  Double-buffered on the sequence parity: a peer can be at most one
// exchange ahead.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "flgpu_reduce_geom.h"

namespace flgpu {
namespace k {

constexpr int kMailWidth = 320;       // doubles per rank slot: >= NSLOTS + nd_of(kMaxMem)
constexpr int kMaxRanks = red::kMaxRanks;

struct Mailbox {
    double data[2][kMaxRanks][kMailWidth];
    unsigned long long flag[2][kMaxRanks];
    unsigned long long error;         // set to the offending sequence number on a wait timeout
};

struct PeerTable { Mailbox *box[kMaxRanks]; };

// One IPC-shared allocation per rank: the mailbox of the host-driven exchanges (exchange_kernel), a second one for
// the exchanges a device-resident line search performs on its own, and that search's sequence counter (the host
// cannot know how many evaluations a search will make, so the counter lives on the device; every rank makes the
// same evaluations, so the counters agree without communication).
struct MailboxPair {
    Mailbox host_driven;
    Mailbox device_search;
    unsigned long long dseq;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Executed by ONE block: store vals[0..count) (count <= blockDim.x) into slot [me] of every rank's mailbox, publish
// `seq`, wait for every rank's slot of this rank's mailbox, return the rank-tree sums in out[0..count).
// timeout_ns: how long a peer may stay silent before this rank records the failure in its mailbox and traps.
__device__ __forceinline__ void mailbox_exchange_block(const PeerTable &peers, int me, int G, unsigned long long seq,
                                                       const double *vals, int count, double *out,
                                                       unsigned long long timeout_ns) {
    const int par = (int)(seq & 1ull), t = threadIdx.x;
    if (t < count) {
        const double v = vals[t];
        for (int r = 0; r < G; r++) peers.box[r]->data[par][me][t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < G) st_release_sys(&peers.box[t]->flag[par][me], seq);
    Mailbox *mine = peers.box[me];
    if (t < G) {
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(&mine->flag[par][t]) < seq) {
            if (global_timer_ns() - t0 > timeout_ns) {
                mine->error = seq;
                __threadfence_system();
                __trap();
            }
        }
    }
    __syncthreads();
    if (t < count) {
        double v[kMaxRanks];
        for (int r = 0; r < G; r++) v[r] = __ldcv(&mine->data[par][r][t]);
        out[t] = red::rank_tree(v, G);
    }
}

// What a device-resident line search written outside the library needs to trade its partial sums inside the kernel
// (filled by flgpu_comm_search_exchange, flgpu.h): the search mailboxes of all ranks, this rank's place, the device
// counter of exchanges made so far (every rank makes the same evaluations, so the counters agree without
// communication), the timeout, and four doubles of device scratch that hand the sums to the other blocks of the grid.
struct SearchExchange {
    PeerTable peers;
    int me, G;
    unsigned long long *dseq;
    unsigned long long timeout_ns;
    double *glob;            // [2 evaluation parities][2]
};

}  // namespace k
}  // namespace flgpu
