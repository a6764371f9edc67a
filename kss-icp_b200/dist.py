"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

Two shard axes exist on this path and a single pair is never split:
  * registration pairs        -> contiguous blocks per rank, no data-path collective (bench.py)
  * rotation hypotheses of one pair (KSS_ICP.hpp:102-118, a serial `for` in the reference)
                              -> round-robin over ranks, ONE all-reduce(MIN) of the fp64 fitness vector,
                                 after which every rank applies the reference's selection rule locally.
`backend` is anything with the single-object methods of kss_icp_b200.Context (middle_align,
rotation_sweep, apply_similarity, icp, apply_transform, nn_metrics)."""
import numpy as np

JUDGE_THRESHOLD = 0.0005       # KSS_ICP.hpp:99


def shard_range(total, world, rank):
    """contiguous block of `total` items owned by `rank`"""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def select_winner(fitness):
    """KSS_ICP.hpp:100-116: Q = 9999; first index with strict ri < Q and ri >= 0 wins; default 0"""
    q, idx = 9999.0, 0
    for i, r in enumerate(fitness):
        if r < q and r >= 0:
            q, idx = r, i
    return idx


def angle_list(step):
    """angleList values per axis index: index * 6.3 / step for the G loop values of `for (a = 0; a < 6.3; a += 6.3 / step)`
    (initRegistrationKSS.hpp:245, 282-284); the same double arithmetic as kss_sweep_angles"""
    out, a, g = [], 0.0, 0
    while a < 6.3:
        out.append(float(g) * 6.3 / float(step))
        a = a + 6.3 / float(step)
        g += 1
    return np.array(out, np.float64)


def register_hypothesis_sharded(backend, sim_s, sim_t, full_s, full_t, rank=0, world=1, allreduce_min=None,
                                step=8.0, max_iter=1000, list_angles=None):
    """list_angles: per-axis angleList table (default: the one `step` defines).
    KSSICP_Registration (KSS_ICP.hpp:86-130) + PCR_QM with the hypothesis ICP runs sharded over ranks.
    allreduce_min(np.ndarray[float64]) -> np.ndarray reduces element-wise MIN over all ranks in place."""
    if list_angles is None:
        list_angles = angle_list(step)
    a7, al = backend.middle_align(sim_s, sim_t)
    sw = backend.rotation_sweep(al, sim_t, step)
    src0 = backend.apply_similarity(sim_s, a7, sw["best_angle"])
    judge = backend.icp(src0, sim_t, max_iter=max_iter)
    used = np.array(sw["best_angle"], np.float64)
    winner, multi, fit = -1, 0, None
    if judge["fitness"] > JUDGE_THRESHOLD:
        multi = 1
        L = len(sw["minima"])
        fit = np.full(L, np.inf, np.float64)
        for l in range(rank, L, world):                                   # this rank's hypotheses
            ang = list_angles[sw["minima"][l]]
            fit[l] = backend.icp(backend.apply_similarity(sim_s, a7, ang), sim_t, max_iter=max_iter)["fitness"]
        if world > 1:
            fit = allreduce_min(fit)                                      # the single collective of this path
        winner = select_winner(fit)
        used = np.array(list_angles[sw["minima"][winner]], np.float64)
    # KSS_ICP.hpp:130 re-runs the ICP of the chosen start on every rank (deterministic, so identical)
    final = backend.icp(backend.apply_similarity(sim_s, a7, used), sim_t, max_iter=max_iter)
    point_align = backend.apply_transform(final["T"], backend.apply_similarity(full_s, a7, used))
    m = backend.nn_metrics(point_align, full_t)
    return dict(align=a7, best_index=sw["best_index"], n_minima=len(sw["minima"]), branch_multi=multi, winner=winner,
                used_angle=used, judge_fitness=judge["fitness"], final_fitness=final["fitness"], T=final["T"],
                mse=m[0], rmse=m[1], mae=m[2], fitness_vector=fit, point_align=point_align)
