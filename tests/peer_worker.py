"""Worker of tests/test_multi_gpu.py (launched with torch.distributed.run, one rank per GPU): the multi-GPU query
through the device-side peer-memory exchange and through the host-driven NCCL exchange must both reproduce the
single-GPU answer bit for bit (matches, bracket record, final sums, p_n) for hits placed on rank boundaries."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["PSS_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pss_b200 import _lib, search
    h = _lib.default_handle()
    n_max = 3_000_000
    cases = []
    # reference answers from the single-GPU path on the same device
    hi = h.nth_prime_upper(n_max) + 1
    slices = search.partition_range(hi, world)
    counts = [h.count_primes(lo, hi_g) for lo, hi_g in slices]
    bounds = [sum(counts[:g]) for g in range(1, world)]          # index of the last prime of ranks 0..world-2
    hits = sorted({1, 7, n_max, n_max - 1} | {b + d for b in bounds for d in (-1, 0, 1, 2)})
    for power, allp in ((2, False), (3, False), (5, True)):
        for hit in hits[:: (1 if power == 2 else 3)]:
            tgt = h.primesum(hit, power)
            for op in ("==", ">="):
                cases.append((tgt, power, allp, op, 1))
        cases.append((10 ** 90, power, allp, "==", 1))            # no match: bracket at n_max
        cases.append((h.primesum(bounds[0] + 1, power), power, allp, "==", bounds[0] + 2))   # hit just below n_min
    bad = 0
    for mode in ("peer", "nccl"):
        os.environ["PSS_EXCHANGE"] = mode
        for tgt, power, allp, op, n_min in cases:
            exp = search.search(tgt, n_max, power=power, n_min=n_min, op=op, all_powers=allp, handle=h)
            got = search.distributed_search(tgt, n_max, power=power, n_min=n_min, op=op, all_powers=allp, handle=h)
            ok = (got.first() == exp.first() and got.last_prime == exp.last_prime and len(got.powers) == len(exp.powers))
            for a, b in zip(got.powers, exp.powers):
                ok = ok and (a.power, a.match_count, a.first_match_n, a.last_match_n, a.cross_n, a.cross_prime, a.sum_at_cross,
                             a.sum_before_cross, a.final_sum) == (b.power, b.match_count, b.first_match_n, b.last_match_n, b.cross_n,
                                                                  b.cross_prime, b.sum_at_cross, b.sum_before_cross, b.final_sum)
            if not ok:
                bad += 1
                print(f"[rank {rank}] MISMATCH mode={mode} power={power} all={allp} op={op} n_min={n_min} first={got.first()} exp={exp.first()}",
                      flush=True)
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"PEER_WORKER cases={len(cases) * 2} mismatches={int(t.item())}", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
