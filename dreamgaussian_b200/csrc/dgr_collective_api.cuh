// dgr_collective_api.cuh — C-ABI entry points of the multi-GPU collectives (include/dgr_b200.h: dgr_peer_*).  Included by dgr_api.cu
// (one translation unit); kept apart from it because none of this runs in a single-GPU forward + backward step.
#pragma once

size_t dgr_peer_flag_bytes(void) { return (size_t)2 * kMaxFlagBlocks * kMaxPeers * sizeof(unsigned); }

int dgr_peer_allreduce(const uint64_t *peer_ptrs, int32_t world, int32_t rank, uint64_t n_floats, uint64_t multicast_ptr,
                       const uint64_t *peer_flag_ptrs, uint32_t epoch, void *stream) {
    NvtxRange nvtx_("dgr_peer_allreduce");
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(-1, "bad world / rank");
    if (n_floats % 4 != 0) return fail(-1, "n_floats must be a multiple of 4");
    if (!peer_ptrs && !multicast_ptr) return fail(-1, "no peer pointers");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n4 = (size_t)(n_floats / 4);
    if (n4 == 0 || world == 1) return 0;
    DevInfo *dv = dev_info();
    if (!dv) return -1;
    const size_t per = (n4 + world - 1) / world;
    int grid = (int)((per + 512 * kUnroll - 1) / (512 * kUnroll));
    if (grid > 4 * dv->sms) grid = 4 * dv->sms;
    if (grid > kMaxFlagBlocks) grid = kMaxFlagBlocks;
    if (grid < 1) grid = 1;
    PeerFlags pf;
    for (int w = 0; w < kMaxPeers; w++) pf.p[w] = (peer_flag_ptrs && w < world) ? reinterpret_cast<unsigned *>(peer_flag_ptrs[w]) : nullptr;
    const bool bar = peer_flag_ptrs != nullptr;
    if (multicast_ptr) {
        float *mc = reinterpret_cast<float *>(multicast_ptr);
        static const bool pipe = !(getenv("DGR_AR_PIPELINE") && getenv("DGR_AR_PIPELINE")[0] == '0');       // A/B switch, read once
        if (bar && pipe) DGR_KERNEL("allreduce_multimem", st, 0, allreduce_multimem_kernel<true, true><<<grid, 512, 0, st>>>(mc, pf, epoch, world, rank, n4));
        else if (bar) DGR_KERNEL("allreduce_multimem", st, 0, allreduce_multimem_kernel<true, false><<<grid, 512, 0, st>>>(mc, pf, epoch, world, rank, n4));
        else if (pipe) DGR_KERNEL("allreduce_multimem", st, 0, allreduce_multimem_kernel<false, true><<<grid, 512, 0, st>>>(mc, pf, epoch, world, rank, n4));
        else DGR_KERNEL("allreduce_multimem", st, 0, allreduce_multimem_kernel<false, false><<<grid, 512, 0, st>>>(mc, pf, epoch, world, rank, n4));
    } else {
        PeerPtrs pp;
        for (int w = 0; w < kMaxPeers; w++) pp.p[w] = w < world ? reinterpret_cast<float *>(peer_ptrs[w]) : nullptr;
        if (bar) DGR_KERNEL("allreduce_p2p", st, 0, allreduce_p2p_kernel<true><<<grid, 512, 0, st>>>(pp, pf, epoch, world, rank, n4));
        else DGR_KERNEL("allreduce_p2p", st, 0, allreduce_p2p_kernel<false><<<grid, 512, 0, st>>>(pp, pf, epoch, world, rank, n4));
    }
    return 0;
}

int dgr_peer_reduce_staged(const uint64_t *peer_ptrs, const DgrPeerPush *push, int64_t P, int32_t n_seg, const int64_t *seg_off,
                           const int32_t *seg_stride, uint64_t stage_ptr, uint64_t padded_floats, uint64_t multicast_ptr,
                           const uint64_t *peer_flag_ptrs, uint32_t epoch, void *stream) {
    NvtxRange nvtx_("dgr_peer_reduce_staged");
    if (!push || !peer_ptrs || !stage_ptr) return fail(-1, "NULL argument");
    if (int e = check_push(push, P)) return e;
    FlatSegs segs;
    if (int e = flat_segs(n_seg, seg_off, seg_stride, &segs)) return e;
    if (P <= 0 || P > 0x7fffffff) return fail(-1, "bad P");
    const int world = push->world, rank = push->rank;
    if (world == 1) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DevInfo *dv = dev_info();
    if (!dv) return -1;
    size_t widest = 1;
    for (int k = 0; k < segs.n; k++) widest = std::max(widest, (size_t)segs.stride[k]);
    const size_t n4 = (size_t)push->gaussians_per_owner * widest / 4;                     // the longest segment loop (every segment uses the whole grid)
    int grid = (int)((n4 + 512 * kUnroll - 1) / (512 * kUnroll));
    if (grid > 4 * dv->sms) grid = 4 * dv->sms;
    if (grid < 1) grid = 1;
    PeerPtrs pp;
    PeerFlags pf;
    for (int w = 0; w < kMaxPeers; w++) {
        pp.p[w] = w < world ? reinterpret_cast<float *>(peer_ptrs[w]) : nullptr;
        pf.p[w] = (peer_flag_ptrs && w < world) ? reinterpret_cast<unsigned *>(peer_flag_ptrs[w]) : nullptr;
    }
    float *mc = reinterpret_cast<float *>(multicast_ptr);
    const float *stage = reinterpret_cast<const float *>(stage_ptr);
    if (peer_flag_ptrs)
        DGR_KERNEL("reduce_staged", st, 0, reduce_staged_kernel<true><<<grid, 512, 0, st>>>(pp, mc, stage, (size_t)padded_floats, pf, epoch, world, rank, (int)P, (int)push->gaussians_per_owner, segs));
    else
        DGR_KERNEL("reduce_staged", st, 0, reduce_staged_kernel<false><<<grid, 512, 0, st>>>(pp, mc, stage, (size_t)padded_floats, pf, epoch, world, rank, (int)P, (int)push->gaussians_per_owner, segs));
    return 0;
}

int dgr_peer_push_flat(const float *local, const DgrPeerPush *push, int64_t P, int32_t n_seg, const int64_t *seg_off,
                       const int32_t *seg_stride, void *stream) {
    NvtxRange nvtx_("dgr_peer_push_flat");
    if (!push || !local) return fail(-1, "NULL argument");
    if (int e = check_push(push, P)) return e;
    FlatSegs segs;
    if (int e = flat_segs(n_seg, seg_off, seg_stride, &segs)) return e;
    if (P <= 0 || P > 0x7fffffff) return fail(-1, "bad P");
    if (push->world == 1) return 0;
    DevInfo *dv = dev_info();
    if (!dv) return -1;
    PeerPush pd;
    memset(&pd, 0, sizeof(pd));
    pd.per = (int)push->gaussians_per_owner;
    for (int w = 0; w < push->world; w++) pd.delta[w] = push->delta_floats[w];
    cudaStream_t st = (cudaStream_t)stream;
    DGR_KERNEL("push_flat", st, 0, push_flat_kernel<<<2 * dv->sms, 512, 0, st>>>(local, pd, push->world, push->rank, (int)P, segs));
    return 0;
}

