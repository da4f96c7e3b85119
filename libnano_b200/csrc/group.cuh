// group.cuh — one process, several B200s: objectives over a dataset whose rows are split across the devices of a
// group context (nb200_init_multi). libnano is ONE process that shards linear::function_t::do_eval over the worker
// threads of its pool (src/linear/function.cpp:53-57, one accumulator per worker, folded by sum_reduce,
// include/nano/core/reduce.h:23-31; one fit per process, src/linear.cpp:30-48); here each device plays the role of
// one worker, driven from the calling host thread:
//
//   x (first device) --peer copy--> every device
//   every device: the pass over ITS rows (the same kernels as a single-device objective)
//   fused groups (single-target kernel, distinct devices with peer access): every launch publishes its reduced
//       column slices straight into the first device's mailbox over NVLink (vgrad_t1.cuh); the first device's stream
//       waits for the ready words and folds the slots in device order;
//   other groups: un-normalised sums per device -> peer copies to the first device -> the same rank-ordered fold.
//
// The summation order is fixed by the partition (device order), so results are bit-identical to the
// one-process-per-GPU path (libnano_b200/sharded.py) on the same partition.
//
// Included by nb200_api.cu after enqueue_eval / enqueue_finish.
#pragma once

namespace nb200
{
// every group evaluation enqueues its per-device launches under this lock: all devices see the evaluations of all
// objectives (clones driven by other host threads, src/machine/tune.cpp:25-42) in ONE order
inline std::mutex g_group_launch_lock;

// contiguous, balanced row blocks (the rule of libnano_b200/sharded.py::shard_rows): the first N % parts hold one more
inline void group_split_rows(const int64_t N, const int parts, std::vector<int64_t>& row0)
{
    row0.assign(static_cast<size_t>(parts) + 1, 0);
    const int64_t base = N / parts, extra = N % parts;
    for (int r = 0; r < parts; ++r)
    {
        row0[static_cast<size_t>(r) + 1] = row0[static_cast<size_t>(r)] + base + (r < extra ? 1 : 0);
    }
}

// Evaluate a group objective. x_dev lives on the first device and is ordered on the first part's stream; gx_out /
// fx_out are addresses the first device can write (its HBM or mapped host memory). Enqueue only.
inline int group_enqueue_eval(nb200_obj* group, const double* x_dev, const bool want_grad, const int64_t n_global,
                              double* gx_out, double* fx_out)
{
    const int      world = static_cast<int>(group->parts.size());
    nb200_obj*     first = group->parts[0];
    const size_t   x_bytes = static_cast<size_t>(group->n) * sizeof(double);
    const int64_t  len     = 1 + group->data->T + group->data->T * group->data->D;

    std::lock_guard<std::mutex> order(g_group_launch_lock);

    // x is replicated: every other device copies it from the first one once the first stream has produced it
    NB200_CUDA_TRY(cudaSetDevice(first->device));
    NB200_CUDA_TRY(cudaEventRecord(group->group_events[0], first->stream));
    for (int r = 1; r < world; ++r)
    {
        nb200_obj* part = group->parts[static_cast<size_t>(r)];
        NB200_CUDA_TRY(cudaSetDevice(part->device));
        NB200_CUDA_TRY(cudaStreamWaitEvent(part->stream, group->group_events[0], 0));
        NB200_CUDA_TRY(cudaMemcpyPeerAsync(part->d_x, part->device, x_dev, first->device, x_bytes, part->stream));
    }

    if (group->group_fused)
    {
        // the first device last: its stream is the one that waits for the others' ready words
        for (int r = world - 1; r >= 0; --r)
        {
            nb200_obj* part = group->parts[static_cast<size_t>(r)];
            NB200_CUDA_TRY(cudaSetDevice(part->device));
            const int status = (r == 0)
                                   ? enqueue_eval(part, x_dev, want_grad, true, n_global, gx_out, fx_out, true)
                                   : enqueue_eval(part, part->d_x, want_grad, true, n_global,
                                                  want_grad ? part->d_out : nullptr, part->d_out + part->n, true);
            if (status != NB200_OK)
            {
                return status;
            }
        }
        return NB200_OK;
    }

    // two-phase: local sums everywhere, gathered on the first device, folded there in device order
    for (int r = world - 1; r >= 0; --r)
    {
        nb200_obj* part = group->parts[static_cast<size_t>(r)];
        NB200_CUDA_TRY(cudaSetDevice(part->device));
        const int status = enqueue_eval(part, r == 0 ? x_dev : part->d_x, want_grad, false, 0, nullptr, nullptr);
        if (status != NB200_OK)
        {
            return status;
        }
        if (r > 0)
        {
            NB200_CUDA_TRY(cudaEventRecord(group->group_events[static_cast<size_t>(r)], part->stream));
        }
    }
    NB200_CUDA_TRY(cudaSetDevice(first->device));
    const size_t sum_bytes = static_cast<size_t>(want_grad ? len : 1 + group->data->T) * sizeof(double);
    NB200_CUDA_TRY(cudaMemcpyAsync(group->group_gather, first->d_reduced, sum_bytes, cudaMemcpyDeviceToDevice,
                                   first->stream));
    for (int r = 1; r < world; ++r)
    {
        nb200_obj* part = group->parts[static_cast<size_t>(r)];
        NB200_CUDA_TRY(cudaStreamWaitEvent(first->stream, group->group_events[static_cast<size_t>(r)], 0));
        NB200_CUDA_TRY(cudaMemcpyPeerAsync(group->group_gather + static_cast<int64_t>(r) * len, first->device,
                                           part->d_reduced, part->device, sum_bytes, first->stream));
    }
    finish_params p{};
    p.reduced     = first->d_reduced;
    p.x           = x_dev;
    p.gx          = want_grad ? gx_out : nullptr;
    p.fx          = fx_out;
    p.D           = group->data->D;
    p.T           = group->data->T;
    p.n_global    = n_global;
    p.flavour     = group->flavour;
    p.l1          = group->l1;
    p.l2          = group->l2;
    p.mail        = group->group_gather;
    p.world       = world;
    p.col_stride  = len;
    p.reduced_out = first->d_reduced;
    const int64_t blocks = want_grad ? (group->n + 255) / 256 : 1;
    finish_kernel<<<static_cast<unsigned>(blocks), 256, 0, first->stream>>>(p);
    NB200_CUDA_TRY(cudaGetLastError());
    first->launches += 1;
    return NB200_OK;
}

// evaluation entry of the device-resident solvers: a plain objective or a group
inline int solver_enqueue_eval(nb200_obj* obj, nb200_obj* group, const double* x_dev, const int64_t n_global,
                               double* gx_out, double* fx_out, const bool want_grad = true)
{
    if (group != nullptr)
    {
        return group_enqueue_eval(group, x_dev, want_grad, n_global, gx_out, fx_out);
    }
    // one launch; with connected peers the cross-rank all-reduce is part of it (vgrad_t1.cuh)
    return enqueue_eval(obj, x_dev, want_grad, true, n_global, want_grad ? gx_out : nullptr, fx_out, obj->peer_world > 1);
}
} // namespace nb200
