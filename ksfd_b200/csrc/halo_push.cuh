// Halo push fused into the kernel that produces a vector (blas1_kernels.cuh: the Krylov
// producers; sweep_op.cuh: the Richardson sweep).  Device functions and structs only.
#pragma once
#include "device_common.cuh"

// ---------------------------------------------------------------------------
// Halo push fused into the PRODUCER of a vector (several ranks, NVLink peer memory).
// The kernel that writes a Krylov vector also stores its two bottom planes into the
// lower neighbour's ghost buffer and its two top planes into the upper neighbour's
// (plain stores over NVLink; in the plane-SoA layout the planes are the first and the
// last `cnt` doubles of the vector), and the block that finishes last publishes the
// exchange number in the neighbours' flag words and advances this rank's counter.
// Nobody waits here: the consumer (the TMA-fed marcher, tma_march.cuh: halo_arrived)
// waits in the CTAs that read ghost planes.  No exchange kernel is launched at all.
// ---------------------------------------------------------------------------
struct HaloPush {
    double *up_lo0, *dn_hi0;                // parity-0 destinations (nullptr: no push)
    long long pstride;                      // doubles between the two parity buffers
    long long cnt;                          // doubles in the two boundary planes
    long long top0;                         // first element of the two top planes (n - cnt)
    volatile unsigned long long *up_flag_lo, *dn_flag_hi;
    unsigned long long *ctr;                // this rank's exchange counter of the slot
    unsigned *done;                         // block counters: [0] combined, [1] bottom planes (publish_dir)
    const volatile unsigned long long *dead;
};
__device__ __forceinline__ bool halo_push_on(const HaloPush &hp)
{
    return hp.up_lo0 != nullptr && !(hp.dead && *hp.dead);
}
// parity shift of the exchange this launch makes; every block reads the counter
// before the last one (which only exists after all blocks have stored) advances it
__device__ __forceinline__ long long halo_push_shift(const HaloPush &hp, unsigned long long &q)
{
    q = *reinterpret_cast<volatile unsigned long long *>(hp.ctr) + 1;
    return (long long)(q & 1ull) * hp.pstride;
}
// returns true when this thread stored into peer memory
__device__ __forceinline__ bool halo_push1(const HaloPush &hp, long long sh, long long e, double v)
{
    bool did = false;
    if (e < hp.cnt) {
        hp.dn_hi0[sh + e] = v;
        did = true;
    }
    if (e >= hp.top0) {
        hp.up_lo0[sh + e - hp.top0] = v;
        did = true;
    }
    return did;
}
// elements 2e, 2e+1 (cnt and top0 are even whenever the double2 path runs)
__device__ __forceinline__ bool halo_push2(const HaloPush &hp, long long sh, long long e2, double2 v)
{
    const long long e = 2 * e2;
    bool did = false;
    if (e < hp.cnt) {
        *reinterpret_cast<double2 *>(hp.dn_hi0 + sh + e) = v;
        did = true;
    }
    if (e >= hp.top0) {
        *reinterpret_cast<double2 *>(hp.up_lo0 + sh + e - hp.top0) = v;
        did = true;
    }
    return did;
}
// Called by every thread of the blocks that stored boundary planes, right after those
// stores and BEFORE the interior stream (`expected` = number of such blocks).
// Hierarchical release: the storing threads do not fence themselves (a system-scope fence
// in a thread of an SM that is streaming stores waits for all of them: measured +2 us per
// producer kernel on top of the rest, profiles/r02_multi_gpu_step_breakdown.txt).  The CTA
// barrier orders every thread's peer stores before thread 0's device-scope fence and its
// (device-scope) arrival on the block counter; the block that arrives last has therefore
// observed all of them and orders its flag stores after them with ONE system-scope fence
// (fences are cumulative in the PTX memory model), so a neighbour that sees the flag sees the
// planes.  Publishing EARLY matters: a kernel whose last action is a store to peer memory
// ends one NVLink round trip later (grid completion waits for the acknowledgement) — measured
// +6-9 us per producer kernel when the flags went out at the end; now the round trip of the
// planes and of the flags hides behind the interior stream of the same kernel.
__device__ __forceinline__ void halo_push_publish(const HaloPush &hp, unsigned long long q,
                                                  unsigned expected)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(hp.done, 1u);
        if (t == expected - 1) {                // last boundary block: all planes are on their way
            atomicExch(hp.done, 0u);
            flag_release(hp.up_flag_lo, q);
            flag_release(hp.dn_flag_hi, q);
            *hp.ctr = q;
        }
    }
}

// One direction at a time (marching kernels whose boundary CTAs finish the top planes long
// before the bottom planes, MarchArgs::rb): dir 0 = the top planes are stored -> the upper
// neighbour's lo flag; dir 1 = the bottom planes -> the lower neighbour's hi flag, and the
// exchange is complete: the counter advances.  Same hierarchical release as above.
__device__ __forceinline__ void halo_push_publish_dir(const HaloPush &hp, unsigned long long q,
                                                      unsigned expected, int dir)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(hp.done + dir, 1u);
        if (t == expected - 1) {
            atomicExch(hp.done + dir, 0u);
            if (dir == 0) {
                flag_release(hp.up_flag_lo, q);
            } else {
                flag_release(hp.dn_flag_hi, q);
                *hp.ctr = q;
            }
        }
    }
}
