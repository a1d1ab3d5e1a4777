// Deferred half of the per-step statistics all-reduce over peer memory (include/gm3d.h: gm3d_step_reduce_t with
// defer = 1, gm3d_step_reduce_collect): the loss launches of a range of step slots only PUSH their {sum, sum_sq,
// count} into every rank's inbox; this one small launch then waits -- once for the whole range -- until every rank's
// push of every slot has arrived, sums them in rank order and writes the (n, 4) head rows.
//
// Replaces misc.all_reduce_mean (/root/reference/Point-MAE_SA3D/util/misc.py:345-353, call sites
// engine_pretrain_Classifier_SVM.py:297-305) for the loss scalars of n steps.
#include "loss_reduce.cuh"

namespace gm3d {

__global__ void __launch_bounds__(1024) step_reduce_collect_kernel(const __grid_constant__ gm3d_step_reduce_t r, int n) {
    __shared__ float s_in[1024][3];
    __shared__ int s_missing;
    const int tid = threadIdx.x;
    if (tid == 0) s_missing = 0;
    __syncthreads();
    const int slot = tid / r.world, src_rank = tid % r.world;
    // Which launch of the slot to sum.  Plain form: the current one (the slot's loss launch precedes this kernel in stream
    // order and has counted itself).  Lagging form: the one after the last summed -- provided this rank has pushed it
    // (the counter only grows, so a racing read can only err on the side of skipping).
    bool active = slot < n;
    unsigned epoch = 0;
    if (active) {
        const unsigned cur = *reinterpret_cast<volatile const unsigned*>(r.epoch + slot);
        epoch = r.collected ? r.collected[slot] + 1u : cur;
        active = epoch != 0u && epoch <= cur;
    }
    __syncthreads();  // every thread of a slot has read collected[slot] before its first thread advances it
    if (active) {
        const InboxSlot* src = reinterpret_cast<const InboxSlot*>(static_cast<const char*>(r.inbox[r.rank]) +
                                                                  static_cast<size_t>(slot) * GM3D_INBOX_BYTES) +
                               (epoch & (GM3D_INBOX_DEPTH - 1u)) * GM3D_MAX_PEERS + src_rank;
        float a, b, c;
        if (!inbox_wait(src, epoch, r.timeout_us, a, b, c)) atomicMax(&s_missing, src_rank + 1);
        s_in[tid][0] = a, s_in[tid][1] = b, s_in[tid][2] = c;
    }
    __syncthreads();
    if (active && src_rank == 0 && r.collected) r.collected[slot] = epoch;
    if (active && src_rank == 0 && r.head) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int q = 0; q < r.world; ++q) {
            a = __fadd_rn(a, s_in[tid + q][0]), b = __fadd_rn(b, s_in[tid + q][1]), c = __fadd_rn(c, s_in[tid + q][2]);
        }
        float* h = r.head + 4 * slot;
        h[0] = a, h[1] = b, h[2] = c, h[3] = static_cast<float>(r.world);
    }
    if (tid == 0 && s_missing && r.status) *r.status = s_missing;
}

}  // namespace gm3d

GM3D_API int gm3d_step_reduce_collect(const gm3d_step_reduce_t* first, int n, void* stream) {
    using namespace gm3d;
    if (!first || n <= 0 || first->world < 1 || first->world > GM3D_MAX_PEERS || !first->epoch || first->rank < 0 ||
        first->rank >= first->world || n * first->world > 1024)
        return GM3D_EINVAL;
    const int threads = (n * first->world + 31) & ~31;
    step_reduce_collect_kernel<<<1, threads, 0, as_stream(stream)>>>(*first, n);
    return launch_status();
}
