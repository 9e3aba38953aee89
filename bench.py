#!/usr/bin/env python
"""
bench.py — headline benchmark of the hot path (BASELINE.json metric):
    primes sieved + power-summed per second, bit-exact, n <= 1e9, p = 2..5, on 1/2/4/8 B200.

A "step" is one complete pass of the hot path over one workload.  The headline workload is
    C3:  does_exist primesum(n,2) == trisum(666), n in [1, 1e9]        (BASELINE.json configs[2])
i.e. sieve [0, 2.28e10) into the wheel-30 bitmap -> exact sums of p^2 per 2880-integer block, warp tile and scan
tile -> grid-level prefix scan -> compare against the 325-bit target (the block whose exact sum interval brackets
the target is walked prime by prime, the block holding n_max is summed by its warp; every other block compares like
its prefix because S(n) is strictly increasing) -> match / bracket record.  Inputs are only the query parameters: nothing
is pre-computed or cached between steps (bitmap and block records are rewritten every step; both are far larger
than L2).  The same JSON line carries, timed with the same --steps / --warmup:
    extra_workloads   C5 (all powers 1..5, n <= 1e9, window n >= 5e7), C4_k3 / C4_k4 / C4_k5 (the wide-limb ranges),
                      C2 (for_any primesum(n,2) == tri(m), n <= 1e8, m <= 1e7; single-GPU API, N = 1 only)
    exhaustive_compare  C3 with pss_set_compare_mode(1): every prime decoded and squared, every S(n) formed and compared
    verified_multi    (N > 1) synthetic-hit queries on every rank boundary +-1, n_max and one '>=' for k = 2 and for all
                      powers 1..5 through distributed_search, compared field by field with the single-GPU answer

  value      primes / device time: CUDA events on the library's launch stream from the first launch to the end of the
             last kernel of each query (gaps and, at N > 1, the waits on the peers included); max over ranks.
  e2e        the same metric through the public Python API (ctypes -> C ABI with host buffers): wall clock around K
             calls, barrier + synchronize on both sides, max over ranks.
  roofline   dominant kernel (k_sieve) against the roof that binds it: shared-memory cross-offs X(N)/t vs
             SMs x 32 banks x f_SM (clock sampled during the timed region); the HBM figure SURVEY 8(d) defines
             (8 B per prime per kernel, 16 B for the path) is kept under hbm_nominal / path_nominal.
  cpu_baseline  oracle port (oracle/oracle.c = restatement of the reference's _segmented_sieve + exact power sums) on
             all host cores, bounded sample, rank 0 at N = 1 only.

--impl reference times that CPU port alone on the same workloads (same keys; the reference itself is pure Python
and O(n^2) on this query; its linear pieces are what oracle.c restates).  When the reference tree is importable
(/root/reference or baseline/_ref; never on the GPU box) its own generate_primes + cpu_power_sum are timed on a
stated prefix as "B0".

Multi-GPU (torchrun, one rank per GPU): the fixed 1e9-prime job is split by value range -> "strong" scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
sys.path.insert(0, ROOT)

TRISUM_666 = 37005443752611483714216385166550857181329086284892731078593232926279977894581784762614450464857290
S1E9 = {1: 11138479445180240497, 2: 168072405102068540986037048787,
        3: 2863851545705647609829761010496201443383,
        4: 52128134804553784574781038633583251938565663646639,
        5: 989096022025420928255871308724946762578387567802486376275087}
WORKLOADS = {
    # name: n_max, power, all_powers, n_min, expected p_n and final sums by power
    "C3": dict(n_max=10 ** 9, power=2, all_powers=False, n_min=1, p_n=22801763489, sums={2: S1E9[2]}),
    "C5": dict(n_max=10 ** 9, power=5, all_powers=True, n_min=5 * 10 ** 7, p_n=22801763489, sums=S1E9),
    "C4_k3": dict(n_max=5 * 10 ** 7, power=3, all_powers=False, n_min=1, p_n=982451653,
                  sums={3: 11387341540580442933005483856330573}),
    "C4_k4": dict(n_max=10 ** 7, power=4, all_powers=False, n_min=1, p_n=179424673,
                  sums={4: 1978202922332179152625028353627144001199}),
    "C4_k5": dict(n_max=5 * 10 ** 6, power=5, all_powers=False, n_min=1, p_n=86028121,
                  sums={5: 3731581465328138542376820998594833918115050077}),
    "C2": dict(n_max=10 ** 8, power=2, all_powers=False, n_min=1, p_n=2038074743, tri_m_max=10 ** 7,
               tri_hits=[{"m": 36, "n": 7}, {"m": 3169, "n": 86}], sums={2: 133759354162117403400944283}),
    "small": dict(n_max=10 ** 6, power=2, all_powers=False, n_min=1, p_n=15485863, sums={2: 76304519151822049179}),
}
EXTRA = ["C5", "C4_k3", "C4_k4", "C4_k5", "C2"]
IMAD_PER_PRIME = {1: 0, 2: 4, 3: 10, 4: 18, 5: 28}   # IMAD.WIDE of the per-prime power chain
MISC_BYTES = 2016   # sizeof(MiscDev): ticket/flag + 8 result records + carry limbs + timeline stamps, one H2D and one D2H per scan
MISC_INIT_BYTES = 424   # sizeof(MiscInit): by-value argument of k_init_misc
PEER_RESULT_BYTES = 1792   # sizeof(PeerResult): one rank's outcome record
SMS = 148


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return {}


_BASE_PRIMES = None


def crossoffs(lo: int, hi: int) -> float:
    """algorithmic wheel-30 cross-offs of a sieve of [lo, hi):
    X = sum over primes 7 <= p <= sqrt(hi) of floor((hi - max(lo, p^2)) / p) * 8/30."""
    global _BASE_PRIMES
    import math
    lim = math.isqrt(max(hi, 4)) + 1
    if _BASE_PRIMES is None or _BASE_PRIMES[-1] < lim:
        import numpy as np
        top = max(lim, 1 << 18)
        s = np.ones(top + 1, dtype=bool)
        s[:2] = False
        for i in range(2, math.isqrt(top) + 1):
            if s[i]:
                s[i * i::i] = False
        _BASE_PRIMES = [int(x) for x in np.nonzero(s)[0]]
    x = 0.0
    for p in _BASE_PRIMES:
        if p < 7:
            continue
        if p * p >= hi:
            break
        x += ((hi - max(lo, p * p)) // p) * 8.0 / 30.0
    return x


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The timed region of this workload is tens of
    milliseconds, far shorter than `nvidia-smi -lms` can resolve, so NVML is polled from a thread every ~2 ms
    (the main thread sits in ctypes calls with the GIL released); nvidia-smi is the fallback."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.samples, self.reason_bits, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def _run(self):
        nv, hdl = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(hdl, nv.NVML_CLOCK_SM))
                try:
                    self.reason_bits |= nv.nvmlDeviceGetCurrentClocksEventReasons(hdl)
                except Exception:
                    self.reason_bits |= nv.nvmlDeviceGetCurrentClocksThrottleReasons(hdl)
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu_index]) if vis and vis.split(",")[self.gpu_index].isdigit() else self.gpu_index
            hdl = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(hdl, nv.NVML_CLOCK_SM)
            self._nvml = (nv, hdl)
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "nvml"}
        if self._thread is None:
            return self._smi_once()
        self._stop.set()
        self._thread.join(timeout=2)
        if self.samples:
            sm = sorted(self.samples)
            out.update(sm_mhz=float(sm[len(sm) // 2]), sm_mhz_min=float(sm[0]), sm_mhz_max_seen=float(sm[-1]), samples=len(sm),
                       reasons=sorted(name for bit, name in self.REASONS.items() if self.reason_bits & bit))
        return out

    def _smi_once(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvidia-smi (single sample after the timed region)"}
        try:
            txt = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = [float(x) for x in txt.strip().split(",")[:2]]
            out.update(sm_mhz=a, sm_max_mhz=b, samples=1)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_baseline(workload, seconds=12.0, threads=None, standin_seconds=0.0):
    """Oracle port on all host cores: sieve + exact power sums over value blocks spread evenly across the
    workload's range, time-bounded.  Returns primes/s and the measured seconds of the pass."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    orc.lib()
    w = WORKLOADS[workload]
    hi = w["p_n"] + 1
    powers = list(range(1, w["power"] + 1)) if w["all_powers"] else [w["power"]]
    cores = threads or len(os.sched_getaffinity(0))
    block = 1 << 24
    n_blocks_total = max(1, hi // block)
    # evenly spaced sample of blocks (same density profile as the full range), processed until time is up
    order = [(i * 7919) % n_blocks_total for i in range(n_blocks_total)]
    lock = threading.Lock()

    def run(fn, budget):
        st = {"next": 0, "primes": 0, "blocks": 0}
        t0 = time.perf_counter()
        deadline = t0 + budget

        def worker():
            while time.perf_counter() < deadline:
                with lock:
                    i = st["next"]
                    st["next"] += 1
                if i >= len(order):
                    return
                b = order[i]
                n = fn(b * block, min((b + 1) * block, hi))
                with lock:
                    st["primes"] += n
                    st["blocks"] += 1

        with ThreadPoolExecutor(cores) as ex:
            for f in [ex.submit(worker) for _ in range(cores)]:
                f.result()
        return st, time.perf_counter() - t0

    st, dt = run(lambda lo, hi_: orc.c_block_sieve_sum(lo, hi_, powers, 1 << 18)[0], seconds)
    whole = st["blocks"] >= len(order)
    res = {
        "value": st["primes"] / dt, "unit": "primes/s", "cores": cores, "kind": "port", "seconds": dt,
        "sample": f"{st['blocks']} of {len(order)} value blocks of 2^24 spread evenly over [0, {hi}) "
                  f"({st['primes']} primes, powers {powers}) in {dt:.2f} s{' = the whole range' if whole else ''}; oracle/oracle.c "
                  "(restatement of utils/sieve.py:_segmented_sieve + exact sum p^k), one block per thread",
    }
    if standin_seconds > 0:
        # the reference's preferred sieve (primesieve) is a third-party C++ library that is not installed here: report a
        # conventional bit-packed wheel-30 segmented sieve + 256-bit sums on the same blocks as a stand-in for it
        st2, dt2 = run(lambda lo, hi_: orc.c_fast_block_sieve_sum(lo, hi_, powers)[0], standin_seconds)
        res["primesieve_standin"] = {
            "value": st2["primes"] / dt2, "unit": "primes/s", "cores": cores,
            "what": "oracle.c:orc_fast_block_sieve_sum -- cache-blocked bit-packed wheel-30 segmented sieve + 256-bit accumulators "
                    f"(not reference code; primesieve itself is not installed): {st2['blocks']} blocks of 2^24, {st2['primes']} primes in {dt2:.1f} s"}
    return res


def reference_b0(limit=20_000_000, power=2):
    """B0: the reference AS SHIPPED — utils.sieve.generate_primes(N, 'segmented') followed by
    utils.gpu.cpu_power_sum (its own multiprocessing.Pool) — timed on a stated prefix when the reference tree is
    importable.  It never is on the GPU box (/root/reference does not travel), and says so."""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "utils")):
            code = (
                "import sys, time, json, os\n"
                f"sys.path.insert(0, {cand!r})\n"
                "os.environ['PRIME_SQUARE_SUM_QUIET'] = '1'\n"
                "import warnings; warnings.simplefilter('ignore')\n"
                "from utils import sieve, gpu\n"
                f"t0 = time.perf_counter(); p = sieve.generate_primes({limit}, algorithm='segmented'); t1 = time.perf_counter()\n"
                f"s = gpu.cpu_power_sum(p, {power}); t2 = time.perf_counter()\n"
                "print(json.dumps({'primes': int(len(p)), 'sum': str(s), 'sieve_s': t1 - t0, 'power_sum_s': t2 - t1}))\n")
            try:
                out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
                d = json.loads(out.stdout.strip().splitlines()[-1])
                tot = d["sieve_s"] + d["power_sum_s"]
                return {"available": True, "tree": cand, "value": d["primes"] / tot, "unit": "primes/s",
                        "cores": len(os.sched_getaffinity(0)),
                        "sample": f"utils.sieve.generate_primes({limit}, 'segmented') {d['sieve_s']:.2f} s + utils.gpu.cpu_power_sum(primes, {power}) "
                                  f"{d['power_sum_s']:.2f} s over the first {d['primes']} primes (prefix of the workload, not the whole range)",
                        "sum": d["sum"]}
            except Exception as e:  # noqa: BLE001
                return {"available": False, "why": f"reference import/run failed: {e}"}
    return {"available": False, "why": "reference tree not present on this box (/root/reference and baseline/_ref absent): "
                                       "B0 is recorded from the build container in profiles/r02/b0_reference_cpu.json"}


def run_reference(args):
    """--impl reference: the CPU port on all host cores, same workloads and keys as the GPU arm.  A step is one
    bounded pass (the whole range if it finishes inside the per-step budget); ms_per_step is its measured time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = args.workload
    n_steps = args.warmup + args.steps
    # the whole run must end within a few minutes: budget per step of the headline workload
    budget = min(args.cpu_seconds, max(1.0, 150.0 / n_steps))
    vals, secs, last = [], [], None
    for i in range(n_steps):
        r = cpu_baseline(w, seconds=budget, standin_seconds=(min(6.0, budget) if i == n_steps - 1 else 0.0))
        if i >= args.warmup:
            vals.append(r["value"])
            secs.append(r["seconds"])
        last = r
    v = sum(vals) / len(vals)
    last["value"] = v
    extra = {}
    if not args.no_extra:
        for name in EXTRA:
            r = cpu_baseline(name, seconds=3.0)
            extra[name] = {"value": r["value"], "unit": "primes/s", "ms_per_step": r["seconds"] * 1e3,
                           "e2e": {"value": r["value"], "unit": "primes/s"}, "cpu_baseline": r,
                           "config": workload_config(name, 1)}
    line = {
        "impl": "reference", "metric": "primes sieved+power-summed/sec (bit-exact)", "value": v, "unit": "primes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 -> 32-bit limbs (exact)",
        "data": "synthetic (the primes themselves; no external data)",
        "config": workload_config(w, args.gpus), "cpu_baseline": last,
        "e2e": {"value": v, "unit": "primes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "baseline_levels": {"port (oracle.c, all cores)": "ran (this line)", "B1 primesieve stand-in": "ran (cpu_baseline.primesieve_standin)",
                            "B0 reference as shipped": reference_b0()},
        "extra_workloads": extra,
    }
    print(json.dumps(line))


def workload_config(name, gpus):
    w = WORKLOADS[name]
    if name == "C2":
        what = f"C2: for_any primesum(n,2) == tri(m), n in [1, {w['n_max']}], m in [1, {w['tri_m_max']}]"
    else:
        what = (f"{name}: does_exist primesum(n,{'p<=' if w['all_powers'] else ''}{w['power']}) == trisum(666), "
                f"n in [{w['n_min']}, {w['n_max']}]")
    return {
        "workload": f"{what} (sieve [0, {w['p_n'] + 1}), {w['n_max']} primes)",
        "n_max": w["n_max"], "power": w["power"], "all_powers": w["all_powers"],
        "partition": f"value range split into {gpus} li(x)-balanced slices" if gpus > 1 else "single slice",
        "l2": "working set rewritten every step (bitmap + block records: 930 MB at n = 1e9, larger than L2; the C4 ranges "
              "are smaller than L2 by definition of the config)",
        "compare": "interval pruning (exact: S_k(n) strictly increasing); the exhaustive per-prime compare is timed "
                   "separately under exhaustive_compare",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exhaustive", action="store_true", help="skip the extra steps with the exhaustive compare")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_workloads (C5, C4, C2)")
    ap.add_argument("--no-verify-multi", action="store_true", help="N > 1: skip the rank-boundary synthetic-hit verification")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3   # timing rule: at least 3 warm-up steps (allocations, pattern build, clocks)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pss_b200 import _lib, search

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["PSS_DEVICE"] = str(local_rank)
    h = _lib.default_handle()
    target = TRISUM_666
    KKEYS = ["ms_sieve", "ms_count_scan", "ms_compact", "ms_power_scan", "ms_bitmap_reduce", "ms_bitmap_compare"]
    TKEYS = ["ms_span", "ms_sieve", "ms_bitmap_reduce", "ms_peer_pre_publish", "ms_peer_wait_totals", "ms_peer_compare_region",
             "ms_bitmap_compare", "ms_peer_wait_results"]     # per-rank timeline of a multi-GPU query (pss_stats)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def make_step(name):
        w = WORKLOADS[name]
        if name == "C2":
            return lambda: search.for_any_tri(w["n_max"], w["tri_m_max"], handle=h)
        if world > 1:
            return lambda: search.distributed_search(target, w["n_max"], power=w["power"], n_min=w["n_min"],
                                                     all_powers=w["all_powers"], handle=h)
        return lambda: search.search(target, w["n_max"], power=w["power"], n_min=w["n_min"], all_powers=w["all_powers"], handle=h)

    def check(name, out):
        """bit-exactness gate (golden: SURVEY Appendix B / tests/golden)"""
        w = WORKLOADS[name]
        if name == "C2":
            assert out == w["tri_hits"], out
            return
        assert out.first() is None and out.bracket() == (w["n_max"], w["n_max"] + 1), "unexpected match"
        assert out.last_prime == w["p_n"], (out.last_prime, w["p_n"])
        for po in out.powers:
            assert po.final_sum == w["sums"][po.power], f"{name}: S_{po.power} mismatch"

    def timed_steps(step, n_steps):
        """K steps bracketed by barrier + synchronize.  The library accumulates its CUDA-event times over the K queries
        (pss_stats since the reset), so the timed loop contains nothing but the K API calls."""
        acc = dict.fromkeys(KKEYS + [k for k in TKEYS if k not in KKEYS], 0.0)
        barrier()
        h.stats_reset()
        t_start = time.perf_counter()
        for _ in range(n_steps):
            o = step()
        barrier()
        wall = time.perf_counter() - t_start
        s = h.stats()
        # whole-query span on the launch stream (gaps and, at N > 1, the waits on the peers' publishes included), summed over
        # the K queries; paths with a host synchronisation in the middle fall back to the sum of the kernels' event times
        dev = s["ms_span"] if s.get("ms_span", 0.0) > 0.0 else s["ms_total_device"]
        for k in acc:
            acc[k] = s.get(k, 0.0)
        return dev, wall * 1e3, acc, s["launches"], o

    def gather_timeline(kern, wall_ms):
        """every rank's per-step averages of the query timeline (rank order); None at N = 1"""
        if world == 1:
            return None
        t = torch.tensor([kern.get(k, 0.0) / args.steps for k in TKEYS] + [wall_ms / args.steps], dtype=torch.float64, device="cuda")
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        rows = []
        for r, p in enumerate(parts):
            v = p.cpu().tolist()
            row = {"rank": r}
            row.update({k: round(x, 4) for k, x in zip(TKEYS + ["ms_wall_e2e"], v)})
            rows.append(row)
        return rows

    def reduce_ranks(dev_ms, wall_ms, kern, launches):
        if world == 1:
            return dev_ms, wall_ms, kern, launches
        t = torch.tensor([dev_ms, wall_ms] + [kern[k] for k in KKEYS] + [float(launches)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        return float(tmax[0]), float(tmax[1]), {k: float(tmax[2 + i]) for i, k in enumerate(KKEYS)}, int(tsum[2 + len(KKEYS)])

    def run_workload(name, sampler=None):
        """warm-up, then the timed region (K steps; per-kernel events at the library's default level 1 = around the sieve
        only, so the region is not stretched by instrumentation), then K more steps with events around EVERY kernel for
        the per-kernel table and the multi-GPU timeline"""
        step = make_step(name)
        for _ in range(args.warmup):
            out = step()
        check(name, out)
        if sampler is not None:
            sampler.start()
        dev_ms, wall_ms, kern, launches, out = timed_steps(step, args.steps)
        clocks = sampler.stop() if sampler is not None else None
        check(name, out)
        h.set_kernel_events(2)
        try:
            step()
            _, wall_d, kern_d, _, out = timed_steps(step, args.steps)
        finally:
            h.set_kernel_events(1)
        check(name, out)
        timeline = gather_timeline(kern_d, wall_d)
        _, _, kern_d, _ = reduce_ranks(0.0, 0.0, kern_d, 0)
        dev_ms, wall_ms, kern, launches = reduce_ranks(dev_ms, wall_ms, kern, launches)
        kern_d["ms_sieve_timed_region"] = kern["ms_sieve"]
        return dict(dev_ms=dev_ms, wall_ms=wall_ms, kern=kern_d, launches=launches, clocks=clocks, timeline=timeline)

    # ---------------------------------------------------------------- N > 1: rank-boundary verification (warm-up)
    verified_multi = None
    if world > 1 and not args.no_verify_multi:
        verified_multi = verify_multi(h, search, dist, torch, world, rank, args.workload)

    # ---------------------------------------------------------------- headline workload
    main_name = args.workload
    sampler = ClockSampler(local_rank) if rank == 0 else None
    res = run_workload(main_name, sampler)
    clocks = res["clocks"]

    # the same steps with the exhaustive per-prime compare (outside the headline timed region)
    exhaustive = None
    if not args.no_exhaustive and main_name != "C2":
        h.set_compare_mode(True)
        h.set_kernel_events(2)
        try:
            step = make_step(main_name)
            step()
            ex_dev, ex_wall, ex_kern, _, ex_out = timed_steps(step, args.steps)
            check(main_name, ex_out)
            ex_dev, ex_wall, ex_kern, _ = reduce_ranks(ex_dev, ex_wall, ex_kern, 0)
            exhaustive = (ex_dev, ex_wall, ex_kern["ms_bitmap_compare"] / args.steps)
        finally:
            h.set_compare_mode(False)
            h.set_kernel_events(1)

    extras = {}
    if not args.no_extra:
        for name in EXTRA:
            if name == main_name or (name == "C2" and world > 1):
                continue
            extras[name] = run_workload(name)

    if os.environ.get("PSS_TRACE") and world > 1:
        print(f"[trace rank {rank}] per-step ms: " + ", ".join(f"{k}={v / args.steps:.3f}" for k, v in res["kern"].items()) +
              f", span={res['dev_ms'] / args.steps:.3f}, wall={res['wall_ms'] / args.steps:.3f}", file=sys.stderr)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    K = args.steps
    peak, peak_src, peaks = load_peaks()
    ncu = load_json("profiles", "ncu_traffic.json")
    micro = load_json("profiles", "microbench_b200.json")
    f_sm = ((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0) * 1e6
    slices = search.partition_range(WORKLOADS[main_name]["p_n"] + 1, world) if world > 1 else None

    def rate(x, ms):
        return x / (ms * 1e-3) if ms else None

    def describe(name, r):
        """value / e2e / per-kernel rates of one workload; every rate is one the kernel really executes"""
        w = WORKLOADS[name]
        n_primes = w["n_max"]
        hi = (w["p_n"] + 1) if world > 1 else None
        per = {k: v / K for k, v in r["kern"].items()}
        per["ms_sieve"] = per.pop("ms_sieve_timed_region", per["ms_sieve"])    # the dominant kernel: timed inside the timed region
        # sieve bound: the proven p_n upper bound the library sieves to (single GPU) / the slices (multi GPU)
        sieve_hi = h.nth_prime_upper(n_primes) + 1
        if world > 1:
            x_alg = max(crossoffs(lo, hi_g) for lo, hi_g in search.partition_range(sieve_hi, world))
            ints = max(hi_g - lo for lo, hi_g in search.partition_range(sieve_hi, world))
        else:
            x_alg = crossoffs(0, sieve_hi)
            ints = sieve_hi
        ms_reduce, ms_compare = per.get("ms_bitmap_reduce", 0.0), per.get("ms_bitmap_compare", 0.0)
        if w["power"] == 2 and not w["all_powers"]:
            reduce_name = "k_bitmap_moments2+k_scan_columns"
        else:
            reduce_name = "k_bitmap_moments_tab+k_scan_columns"
        bitmap_bytes = ints / 30.0
        smem_peak = SMS * 32 * f_sm
        kernels = {
            "k_sieve": {"ms": per["ms_sieve"], "crossoffs_per_s": rate(x_alg, per["ms_sieve"]), "smem_peak_per_s": smem_peak,
                        "smem_frac": (rate(x_alg, per["ms_sieve"]) or 0.0) / smem_peak,
                        "integers_per_s": rate(ints, per["ms_sieve"]), "bitmap_write_GBps": (rate(bitmap_bytes, per["ms_sieve"]) or 0.0) / 1e9},
            reduce_name: {"ms": ms_reduce, "bitmap_read_GBps": (rate(bitmap_bytes, ms_reduce) or 0.0) / 1e9,
                          "primes_per_s": rate(n_primes / world, ms_reduce),
                          "note": "byte-table moments: work is per bitmap word, not per prime; issue-bound (see profiles/ ncu summary)"},
            "k_bitmap_scan<compare>": {"ms": ms_compare, "mode": "interval pruning: only bracketing / n_max / window-edge blocks are walked"},
        }
        if per.get("ms_compact", 0.0):        # materialised path (PSS_SEARCH_PATH=materialised): u64 primes in HBM
            ms_mat = max(per["ms_power_scan"] - ms_reduce - ms_compare, 0.0)
            kernels["k_scan_tile_counts"] = {"ms": per["ms_count_scan"]}
            kernels["k_compact_lut"] = {"ms": per["ms_compact"], "hbm_GBps": (rate(bitmap_bytes + 8.0 * n_primes / world, per["ms_compact"]) or 0.0) / 1e9,
                                        "hbm_peak_GBps": load_peaks()[0]}
            kernels["k_power_tiles+k_scan_columns+k_power_walk"] = {
                "ms": ms_mat, "hbm_GBps_read_once": (rate(8.0 * n_primes / world, ms_mat) or 0.0) / 1e9, "hbm_peak_GBps": load_peaks()[0],
                "note": "8 B per prime read once by pass A; pass B re-reads only the tiles that can hold the target (interval pruning)"}
        dev_ms_total = r["dev_ms"] if r["dev_ms"] > 0 else r["wall_ms"]      # PSS_KERNEL_EVENTS=0 on a multi-call path: wall clock
        dev_step = dev_ms_total / K
        d = {
            "value": n_primes * K / (dev_ms_total * 1e-3), "unit": "primes/s", "ms_per_step": dev_step,
            "e2e": {"value": n_primes * K / (r["wall_ms"] * 1e-3), "unit": "primes/s", "ms_per_step": r["wall_ms"] / K},
            "gpu_launches": r["launches"], "kernels": kernels, "config": workload_config(name, world),
            "kernel_timing": "k_sieve: CUDA events inside the timed region; the other kernels: the same steps repeated with events "
                             "around every kernel (pss_set_kernel_events(2)), which stretches a query by ~20 us",
            "crossoffs_per_prime": x_alg * (world if world > 1 else 1) / n_primes,
            "verified": "final sums, p_n and no-match/bracket checked against tests/golden (bit-exact)" if name != "C2"
                        else "tri hits == [(m=36,n=7),(m=3169,n=86)] (reference find_matches answer)",
        }
        return d, x_alg, reduce_name

    main_d, x_alg, reduce_name = describe(main_name, res)
    n_primes = WORKLOADS[main_name]["n_max"]
    per_sieve_ms = main_d["kernels"]["k_sieve"]["ms"]
    dominant = max(main_d["kernels"], key=lambda k: main_d["kernels"][k].get("ms", 0.0))
    traffic = (ncu.get(dominant) or {}).get("dram_bytes_per_launch")
    sieve_ctr = ncu.get("k_sieve_counters", {})
    hbm_ach = 8 * (n_primes / world) / (per_sieve_ms * 1e-3) / 1e9 if per_sieve_ms else None
    path_ach = 16 * n_primes / (main_d["ms_per_step"] * 1e-3) / 1e9
    roofline = {
        "bound": "smem", "kernel": dominant,
        "achieved": main_d["kernels"]["k_sieve"]["crossoffs_per_s"], "peak": SMS * 32 * f_sm,
        "unit": "shared-memory cross-offs/s (one 4-byte bank access each)",
        "frac": main_d["kernels"]["k_sieve"]["smem_frac"],
        "traffic": traffic if world == 1 else None,
        "peak_source": f"{SMS} SMs x 32 banks x f_SM, f_SM = {f_sm / 1e6:.0f} MHz sampled during the timed region"
                       + (f"; microbenchmark (tools/microbench.cu): conflict-free ATOMS.AND = "
                          f"{micro['atoms_and_conflict_free']['lane_ops_per_s']:.3e} lane-ops/s" if micro.get("atoms_and_conflict_free") else ""),
        "algorithmic_units": f"X(N) = {x_alg:.4e} wheel-30 cross-offs per launch (SURVEY 8(d) formula; "
                             f"{main_d['crossoffs_per_prime']:.2f} per prime)",
        "measured_counters": sieve_ctr or None,
        "hbm_nominal": {"bound": "hbm", "achieved": hbm_ach, "peak": peak, "unit": "GB/s", "frac": (hbm_ach / peak) if hbm_ach else None,
                        "peak_source": peak_src,
                        "note": "SURVEY 8(d) accounting: 8 B per prime per kernel (u64 written once by K1); the fused path really moves "
                                "0.76 B/prime (bit-packed wheel-30 bitmap, see traffic), so HBM does not bind this kernel"},
        "path_nominal": {"achieved_GBps": path_ach, "frac_per_gpu": path_ach / peak / world,
                         "note": "16 B per prime for the path (8 written by K1 + 8 read by K3) over the device span, per GPU"},
        "kernels": main_d["kernels"],
    }
    line = {
        "metric": "primes sieved+power-summed/sec (bit-exact)", "value": main_d["value"], "unit": "primes/s", "n_gpus": world,
        "steps": K, "warmup": args.warmup, "ms_per_step": main_d["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64 -> 32-bit limbs (exact integer)",
        "data": "synthetic (the primes themselves; no external data)",
        "config": workload_config(main_name, world),
        "clocks": clocks,
        "e2e": {"value": main_d["e2e"]["value"], "unit": "primes/s", "ms_per_step": main_d["e2e"]["ms_per_step"],
                # host -> device: the query travels as by-value kernel arguments (result-record init + carry limbs of k_init_misc,
                # target limbs / window / operator of the compare kernel); device -> host: the result block, the column totals
                # and (N > 1) every rank's outcome record, stored by k_export into mapped pinned host memory
                "h2d_bytes_per_step": MISC_INIT_BYTES + 4 * 12 + 40,
                "d2h_bytes_per_step": MISC_BYTES + 8 * (1 + 4) + (PEER_RESULT_BYTES * world if world > 1 else 0),
                "transport": "kernel arguments in, zero-copy stores into mapped pinned memory out (no copy-engine operation in the chain)",
                "api": "pss_b200.search.search -> ctypes -> pss_search (host args, host results)"},
        "gpu_launches": main_d["gpu_launches"],
        "roofline": roofline,
        "verified": main_d["verified"],
    }
    if verified_multi is not None:
        line["verified_multi"] = verified_multi
    if res.get("timeline"):
        line["timeline_per_rank"] = {
            "rows": res["timeline"],
            "legend": "per-step averages on each rank (ms): ms_span = first launch -> every rank's outcome gathered; "
                      "ms_peer_pre_publish = chain start -> slice totals published (ms_sieve + ms_bitmap_reduce + launch gaps); "
                      "ms_peer_wait_totals = spin in k_peer_gather on the lower ranks; ms_peer_compare_region = gather exit -> outcome "
                      "published (ms_bitmap_compare + gaps); ms_peer_wait_results = spin in k_peer_gather_results on all ranks "
                      "(= the other ranks' lateness: launch skew + slower slices); ms_wall_e2e - ms_span = host side of the call"}
    if exhaustive:
        line["exhaustive_compare"] = {
            "value": n_primes * K / (exhaustive[0] * 1e-3), "unit": "primes/s", "ms_per_step": exhaustive[0] / K,
            "e2e_value": n_primes * K / (exhaustive[1] * 1e-3), "e2e_ms_per_step": exhaustive[1] / K,
            "k_bitmap_scan<compare>_ms": exhaustive[2],
            "imad_per_s": IMAD_PER_PRIME[WORKLOADS[main_name]["power"]] * (n_primes / world) / (exhaustive[2] * 1e-3) if exhaustive[2] else None,
            "imad_peak_per_s": (micro.get("imad_wide") or {}).get("thread_instr_per_s"),
            "note": "same query, pss_set_compare_mode(1): every prime decoded and raised to the k-th power, every S_k(n) formed and "
                    "compared (the reference's per-n loop); identical results"}
    if extras:
        line["extra_workloads"] = {}
        for name, r in extras.items():
            d, _, _ = describe(name, r)
            line["extra_workloads"][name] = d
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(main_name, seconds=args.cpu_seconds, standin_seconds=min(6.0, args.cpu_seconds / 2))
        line["cpu_baseline"]["B0_reference_as_shipped"] = reference_b0()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def verify_multi(h, search, dist, torch, world, rank, workload):
    """Multi-GPU parity inside the bench run: synthetic-hit queries (target := S_k(hit) from the single-GPU path) with the
    hit on every rank boundary +-1, at n_max and with one '>=', for k = 2 and for all powers 1..5, through
    distributed_search (the device-side peer exchange) — compared field by field with the single-GPU answer computed
    on the same rank.  Any mismatch fails the run."""
    w = WORKLOADS[workload if workload in ("C3", "C5", "small") else "C3"]
    n_max = w["n_max"]
    hi = h.nth_prime_upper(n_max) + 1
    slices = search.partition_range(hi, world)
    counts = [h.count_primes(lo, hi_g) for lo, hi_g in slices]
    bounds = [sum(counts[:g]) for g in range(1, world)]       # index of the last prime of ranks 0 .. world-2
    hits = sorted({n_max} | {b + d for b in bounds for d in (0, 1, 2) if 1 <= b + d <= n_max})
    cases = []
    for power, allp in ((2, False), (5, True)):
        for i, hit in enumerate(hits):
            # S_k(hit): final sum of a single-GPU query that ends at hit (k = the searched power for the single-power
            # config; for all powers rotate the hit's power)
            pw = power if not allp else 1 + (i % power)
            probe = search.search(0, hit, power=pw, op="<", handle=h)
            cases.append((probe.powers[0].final_sum, power, allp, "==", 1))
        probe = search.search(0, bounds[0] + 1, power=2 if not allp else 3, op="<", handle=h)
        cases.append((probe.powers[0].final_sum, power, allp, ">=", 1))
        cases.append((probe.powers[0].final_sum, power, allp, "==", bounds[0] + 2))      # hit just below n_min
    bad = 0
    for tgt, power, allp, op, n_min in cases:
        exp = search.search(tgt, n_max, power=power, n_min=n_min, op=op, all_powers=allp, handle=h)
        got = search.distributed_search(tgt, n_max, power=power, n_min=n_min, op=op, all_powers=allp, handle=h)
        ok = got.first() == exp.first() and got.last_prime == exp.last_prime and len(got.powers) == len(exp.powers)
        for a, b in zip(got.powers, exp.powers):
            ok = ok and (a.power, a.match_count, a.first_match_n, a.last_match_n, a.cross_n, a.cross_prime, a.sum_at_cross,
                         a.sum_before_cross, a.final_sum) == (b.power, b.match_count, b.first_match_n, b.last_match_n, b.cross_n,
                                                              b.cross_prime, b.sum_at_cross, b.sum_before_cross, b.final_sum)
        if op == "==" and n_min == 1 and exp.first() is None:
            ok = False                       # a synthetic hit must be found
        if not ok:
            bad += 1
            print(f"[verify_multi rank {rank}] MISMATCH power={power} all={allp} op={op} n_min={n_min} got={got.first()} exp={exp.first()}",
                  file=sys.stderr, flush=True)
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    total_bad = int(t.item())
    if rank == 0:
        print(f"[verify_multi] world={world} cases={len(cases)} mismatches={total_bad} rank_boundaries={bounds}", file=sys.stderr, flush=True)
    assert total_bad == 0, f"multi-GPU verification failed: {total_bad} mismatching (rank, case) pairs"
    return {"cases": len(cases), "mismatches": total_bad, "ranks": world,
            "what": "hits on every rank boundary (last prime of a rank, first and second of the next), n_max, one '>=' and one hit "
                    "below n_min; k = 2 and all powers 1..5; distributed_search (peer exchange) vs single-GPU search, all record fields"}


if __name__ == "__main__":
    main()
