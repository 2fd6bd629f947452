"""Multi-GPU plumbing: one process per GPU, torch.distributed for the collective.

Independent policies (BASELINE.json configs 1, 2, 4) shard envs by rank and need NO data-path
collective: rank r owns global envs [r*B, (r+1)*B) (config.env_index0), exactly like r*B..
separate reference processes.

Shared policy (config 3; reference analogue: threads sharing one rl::Agent*, src/main.cpp:196-206)
has exactly one exchange per training tick: a SUM all-reduce of the accumulated weight update
dtheta (memory_size fp64), after which every rank applies theta += dtheta to its replica.
"""


def shard(n_envs_total, rank, world):
    """Contiguous block partition: (first global env index, number of local envs)."""
    base, rem = divmod(n_envs_total, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def shared_policy_tick(accumulate, dtheta, apply, dist=None):
    """One training tick of a policy shared across ranks.

    accumulate(): env tick + learner steps under theta_t, update summed into `dtheta` (a tensor view)
    apply():      theta += dtheta; dtheta = 0; next actions under theta_{t+1}
    dist:         torch.distributed (initialised) or None for a single process
    """
    accumulate()
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(dtheta, op=dist.ReduceOp.SUM)
    apply()


def run_shared_policy(market, n_ticks, dist=None):
    """market: rl_markets_b200.lib.BatchedMarket created with shared_policy=True.

    Device-ordered: the handle is put on torch's CURRENT stream, so per tick the accumulate kernels, the NCCL all-reduce
    of dtheta and the apply kernels are ordered by the stream itself -- the host enqueues all n_ticks ticks without
    waiting for the device once (torch's NCCL work.wait() blocks the stream, not the host)."""
    dth = market.dtheta_tensor() if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
    if dth is None:
        market.run_ticks(n_ticks)
        return
    import torch
    caller = torch.cuda.current_stream()
    if caller.cuda_stream == 0:
        # the legacy default stream cannot be handed to the library (0 means "your own stream" in rlm_set_stream):
        # run the ticks on a side stream that is ordered after, and joined back into, the caller's stream
        side = getattr(market, "_side_stream", None)
        if side is None:
            side = market._side_stream = torch.cuda.Stream()
        side.wait_stream(caller)
        with torch.cuda.stream(side):
            _shared_ticks_on_current_stream(market, n_ticks, dist, dth, torch)
        caller.wait_stream(side)
    else:
        _shared_ticks_on_current_stream(market, n_ticks, dist, dth, torch)


def _shared_ticks_on_current_stream(market, n_ticks, dist, dth, torch):
    cur = torch.cuda.current_stream().cuda_stream
    if getattr(market, "_stream_ptr", None) != cur:
        market.set_stream(cur)
    for _ in range(n_ticks):
        market.shared_tick_accumulate()
        dist.all_reduce(dth, op=dist.ReduceOp.SUM)
        market.apply_dtheta()
