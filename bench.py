#!/usr/bin/env python
"""bench.py -- ELBO-gradient datapoints/s of the AEVB step on synthetic MNIST-shaped data.

Contract (driver): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line from
rank 0.  Workload at every N is BASELINE.json configs[1] ("c2": VAEB.py discrete MNIST 784-d
Bernoulli, Nz=20, 500 tanh hidden, M=100, L=1, Adagrad).  A step = one `update()` = forward +
bound + backward + prior + Adagrad on one minibatch of 100 rows.  M=100 training does not shard
("replicas only"), so N>1 runs N independent replicas (weak scaling, no data-path
collective); the data-parallel config (c3, NCCL all-reduce) and the importance-sampling
estimator (c5) are measured in the same run and reported under "also".

  value  device-resident: x_train already in HBM, K updates enqueued through
         vaeb_update_many (one call, no host sync inside), CUDA events on the launch stream.
  e2e    the reference-facing call with HOST inputs: every step copies its minibatch from
         pinned host memory (H2D), runs the update and reads the bound back (D2H), synchronously.
  --impl reference  the reference's CPU path: it cannot run here (Python 2 + Theano), so this is
         the numpy restatement in oracle/ (kind "port"), fp32, all BLAS threads of the host.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, H, Z, M, L = 784, 500, 20, 100, 1
N_TRAIN = 50000                      # rows resident in HBM: 157 MB > the 126 MB L2
FLOPS_PER_DATAPOINT = 4100000        # SURVEY.md 8d (fwd 1,628,000 + bwd 2,472,000)
METRIC = "ELBO-gradient datapoints/sec (AEVB update, MNIST 784-500-20 Bernoulli, M=100, L=1)"
UNIT = "datapoints/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(object):
    """SM clock and throttle reasons sampled through NVML DURING the timed region (a thread in
    this process: spawning nvidia-smi in a loop contends with kernel launches for the driver)."""

    def __init__(self, index, period_s=0.05):
        self.index, self.period, self.rows, self.stop = index, period_s, [], threading.Event()
        self.max_mhz, self.t = None, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            # the first NVML queries of a process take tens of ms inside the driver and serialise with a kernel
            # launch issued meanwhile (measured: +25..300 ms between the start event and the launch of the timed
            # kernel): pay for them here, before the timed region, and let the thread sleep before its first sample
            self._sample()
            self.rows = []
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception:
            self.t = None
        return self

    def _sample(self):
        nv = self.nv
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), int(reasons), time.perf_counter()))

    def _loop(self):
        while not self.stop.wait(self.period if self.rows else 0.02):
            try:
                self._sample()
            except Exception:
                pass

    def __exit__(self, *a):
        if self.t:
            try:
                self._sample()
            except Exception:
                pass
            self.stop.set()
            self.t.join(timeout=2)

    def summary(self, t0=None, t1=None):
        """Clocks / reasons of the samples taken in [t0, t1] (perf_counter), all samples if no window is given."""
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}
        rows = [r for r in self.rows if (t0 is None or r[2] >= t0) and (t1 is None or r[2] <= t1)]
        if not rows:
            rows = self.rows[-1:]
        sm = [r[0] for r in rows]
        reasons = sorted({n for r in rows for bit, n in names.items() if r[1] & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm)}


def make_problem(seed=15485863):
    from vaeb_b200.data import synthetic_mnist
    return synthetic_mnist(N_TRAIN, seed=seed)


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------
def cpu_step_loop(x, steps, warmup, budget_s=None):
    """fp32 numpy restatement (oracle/vaeb_oracle.py) of the same update on the same shapes."""
    from oracle import vaeb_oracle as O
    m = O.OracleVAEB(x, False, H, Z, M, L=L, params=O.init_params(D, H, Z, False), dtype=np.float32)
    rng = np.random.RandomState(10)
    nb = x.shape[0] // M
    order = np.random.RandomState(1).permutation(nb)
    for i in range(warmup):
        m.update(int(order[i % nb]), rng.normal(size=(L, M, Z)).astype(np.float32))
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        # eps is drawn on the host inside the step, as the reference does (VAEB.py:42)
        m.update(int(order[(warmup + i) % nb]), rng.normal(size=(L, M, Z)).astype(np.float32))
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s and done >= 20:
            break
    dt = time.perf_counter() - t0
    return done, dt


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    x = make_problem()[:5000]          # the CPU arm cycles through 50 minibatches of the same data
    done, dt = cpu_step_loop(x, args.steps, args.warmup)
    val = done * M / dt
    cores = blas_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c2: MNIST-shape Bernoulli VAE D=784 H=500 Z=20, M=100, L=1, Adagrad; one step = "
                                   "one update() on 100 rows", "host_cpus": os.cpu_count()},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d update() steps of the numpy fp32 restatement (oracle/vaeb_oracle.py); the "
                                       "reference itself needs Python 2 + Theano and cannot run here" % done},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import vaeb_b200
    x = make_problem()
    model = vaeb_b200.VAEB(x, False, H, Z, M, L, 0.01, False, False, device=local, seed=10 + rank,
                           precision=args.precision)
    stream = torch.cuda.current_stream()
    model.set_stream(stream.cuda_stream)
    nb = N_TRAIN // M
    rng = np.random.RandomState(15485863 + rank)

    def order(n):
        out = []
        while len(out) < n:
            out += list(rng.permutation(nb))       # np.random.shuffle(batch_order) per epoch (VAEB.py:574)
        return np.asarray(out[:n], dtype=np.int32)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K, W = args.steps, max(args.warmup, 3)
    # ---- value: device-resident, no host sync inside the timed region -------------------
    # the first call also sizes the per-launch buffers (batch order, bounds, pinned readback) for K updates: growing
    # them inside the timed region (cudaFree / cudaMallocHost between the start event and the launch) cost 25-300 ms
    # in some runs -- the spread of `value` seen earlier in the round
    model.update_many(order(max(W, K)))
    # GPU clocks ramp up lazily: keep the device busy for ~1 s before timing (measured: the first 3000
    # updates after a cold start run 10-70% slower than steady state on this pool's B200s)
    # Warm-up continues until two consecutive 1000-update probes agree within 3 % (at least 1 s, at most 6 s).
    # The NVML sampler runs from BEFORE the warm-up to after the timed region and only its samples inside the region
    # are reported, so that none of its start-up work (nvmlInit, first queries) falls into the region.
    clocks = ClockSampler(local)
    clocks.__enter__()
    t_ramp = time.perf_counter()
    probe, last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prev_ms = None
    while True:
        probe.record(stream)
        model.update_many(order(1000))
        last.record(stream)
        torch.cuda.synchronize()
        cur_ms = probe.elapsed_time(last)
        el = time.perf_counter() - t_ramp
        if (el >= 1.0 and prev_ms is not None and abs(cur_ms - prev_ms) <= 0.03 * prev_ms) or el >= 6.0:
            break
        prev_ms = cur_ms
    probe_ms_per_step = cur_ms / 1000.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_region():
        order_k = order(K)
        barrier()
        l0 = model.launch_count()
        t0 = time.perf_counter()
        e0.record(stream)
        out = model.update_many(order_k)
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), model.launch_count() - l0, out, (t0, time.perf_counter())

    ms, launches, elbos, window = timed_region()
    remeasured = None
    if ms / K > 1.3 * probe_ms_per_step:
        # a one-off stall between the start event and the launch: the K steps are timed once more and BOTH results are
        # reported -- `value` is the second one
        remeasured = {"first_ms_per_step": ms / K, "warmup_probe_ms_per_step": probe_ms_per_step}
        ms, launches, elbos, window = timed_region()
    clocks.__exit__(None, None, None)
    clocks_summary = clocks.summary(*window)
    value = world * K * M / (ms * 1e-3)

    # ---- e2e: host minibatch in, bound out, every step --------------------------------
    pinned = torch.empty((N_TRAIN, D), dtype=torch.float32, pin_memory=True)
    pinned.copy_(torch.from_numpy(x))
    xp = pinned.numpy()
    Ke = min(K, 2000)
    oe = order(W + Ke)
    for b in oe[:W]:
        model.update_host(xp[b * M:(b + 1) * M])
    barrier()
    e0.record(stream)
    for b in oe[W:]:
        model.update_host(xp[b * M:(b + 1) * M])
    e1.record(stream)
    barrier()
    ms_e2e_sync = max_over_ranks(e0.elapsed_time(e1))
    e2e_sync = world * Ke * M / (ms_e2e_sync * 1e-3)
    # streaming form (the call a host-side data loader makes): every step still copies ITS minibatch
    # H2D from pinned memory and reads ITS bound back D2H, but the copy of step i+1 overlaps the
    # kernel of step i and the host only waits in collect()
    Ka = K
    oa = order(W + Ka)
    for b in oa[:W]:
        model.update_host_async(xp[b * M:(b + 1) * M])
    model.collect()
    barrier()
    e0.record(stream)
    got = 0
    for i, b in enumerate(oa[W:]):
        model.update_host_async(xp[b * M:(b + 1) * M])
        if (i + 1) % 4096 == 0:
            got += len(model.collect())
    got += len(model.collect())
    e1.record(stream)
    barrier()
    assert got == Ka
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = world * Ka * M / (ms_e2e * 1e-3)

    # ---- also: the two configurations that shard, measured in the same run -------------
    def timed(fn):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        barrier()
        return max_over_ranks(a.elapsed_time(b))

    also = {}
    if not args.no_also:
        from vaeb_b200.data import synthetic_mnist
        from vaeb_b200 import distributed as vd
        peaks_ = measured_peaks()
        # c3: large-batch data-parallel training, global M = 16384 rows, NCCL sum all-reduce of the gradients
        MG = 16384
        per = MG // world
        for prec in ("bf16x3", "bf16"):
            xs = synthetic_mnist(per * 3, seed=777 + rank)
            m3 = vaeb_b200.VAEB(xs, False, H, Z, per, 1, 0.01, False, False, device=local, precision=prec, seed=10)
            m3.set_stream(stream.cuda_stream)
            if world > 1:
                vd.attach_data_parallel(m3)
            m3.update_many(np.arange(3, dtype=np.int32) % 3)
            k3 = 30
            ms3 = timed(lambda: m3.update_many(np.arange(k3, dtype=np.int32) % 3))
            dps = MG * k3 / (ms3 * 1e-3)
            tf = dps * FLOPS_PER_DATAPOINT / 1e12
            also["c3_dp_" + prec] = {
                "workload": "c3: MNIST Bernoulli 784-500-20, global M=16384 (%d rows/GPU), L=1, Adagrad, NCCL all-reduce "
                            "of the 3.26 MB gradient" % per,
                "value": dps, "unit": "datapoints/s", "ms_per_step": ms3 / k3, "steps": k3, "precision": prec,
                "algorithmic_tflops": tf, "frac_of_bf16_sustained_peak_per_gpu": tf / world / peaks_["bf16_tflops_sustained"]}
            m3.close()
        # c5: importance-sampled log p(x), 10,000 test points x L = 5000 samples, points sharded over the ranks.
        # Tensor-core estimator (is_tc.cu, bf16 operands, 1e-2 tier) on the full configuration; the fp32
        # estimator (1e-4 tier) on a 1000-point sample.  Host x in, host log p out: the timed region holds the
        # H2D copy, the fp32 encoder, the fused decoder/log-likelihood/logsumexp kernel and the D2H copy.
        L5 = 5000
        for prec, n_pts in (("bf16", 10000), ("fp32", 1000)):
            xt = synthetic_mnist(n_pts, seed=4242)
            m5 = vaeb_b200.VAEB(xt[:100], False, H, Z, 100, 1, 0.01, False, False, device=local, seed=10, precision=prec)
            m5.set_stream(stream.cuda_stream)
            vd.sharded_log_px(m5, xt, L5, rank, world, gather=False)          # warm-up at the timed size
            res5 = {}
            ms5 = timed(lambda: res5.update(lp=vd.sharded_log_px(m5, xt, L5, rank, world, gather=False)))
            sps = n_pts * L5 / (ms5 * 1e-3)
            tf5 = sps * 804000 / 1e12
            also["c5_is_logpx_" + prec] = {
                "workload": "c5: IS log p(x), %d MNIST-shape points x L=%d, D=784 H=500 Z=20, points sharded over GPUs, "
                            "host x in / host log p out" % (n_pts, L5),
                "value": sps, "unit": "samples/s", "ms": ms5, "precision": prec,
                "algorithmic_tflops": tf5, "flops_per_sample": 804000,
                "frac_of_bf16_sustained_peak_per_gpu": tf5 / world / peaks_["bf16_tflops_sustained"],
                "mean_logpx_rank0": float(np.mean(res5["lp"]))}
            m5.close()
        # c4: full variational Bayes over the weights (VAEB.py --full_varational, getFVBL), discrete MNIST Nz = 2 / 10,
        # M = 100: the reference-faithful mode (weights not sampled) and the sampled-weights mode of the
        # north star (theta = mu + |sigma| zeta per minibatch).  Replica per GPU.
        for zz in (2, 10):
            m0 = vaeb_b200.VAEB(x[:200], False, H, zz, M, 1, 0.01, False, False, device=local, seed=10)
            p0 = m0.get_params()          # the reference initialisation (VAEB.py:50-115) as the MAP start of :117-125
            m0.close()
            for sampled in (False, True):
                m4 = vaeb_b200.VAEB(x[:5000], False, H, zz, M, 1, 0.01, False, True, p0, device=local, seed=10,
                                    sample_weights=sampled)
                m4.set_stream(stream.cuda_stream)
                m4.update_many(np.arange(50, dtype=np.int32) % 50)
                k4 = 1000
                ms4 = timed(lambda: m4.update_many(np.arange(k4, dtype=np.int32) % 50))
                also["c4_fvb_z%d_%s" % (zz, "sampled" if sampled else "faithful")] = {
                    "workload": "c4: full-VB Bernoulli MNIST 784-500-%d, M=100, L=1, %s" % (
                        zz, "weights sampled per minibatch" if sampled else "reference-faithful (weights not sampled)"),
                    "value": world * M * k4 / (ms4 * 1e-3), "unit": "datapoints/s", "ms_per_step": ms4 / k4}
                m4.close()
        # flat Adagrad pass against the HBM roofline (on a buffer far larger than L2 -- at the real parameter
        # counts the 3.3 MB buffers never leave L2): a handle with a 40000-unit hidden layer, 65 M parameters, 1.3 GB
        mo = vaeb_b200.VAEB(x[:200], False, 40000, Z, M, 1, 0.01, False, False, device=local, seed=10)
        mo.set_stream(stream.cuda_stream)
        ms_o, by_o = mo.profile_optimizer(iters=50)
        also["adagrad_flat_stream"] = {
            "workload": "flat Adagrad + prior over 65.2 M parameters (20 B/parameter, 1.30 GB per launch)",
            "value": by_o / (ms_o * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_o,
            "frac_of_measured_hbm_peak": by_o / (ms_o * 1e-3) / 1e9 / peaks_["hbm_gbs"]}
        mo.close()
        # c1: the reference's own CPU-runnable case on one GPU
        from vaeb_b200.data import synthetic_frey
        xf = synthetic_frey()[:1500]
        m1 = vaeb_b200.VAEB(xf, True, 200, 2, 100, 1, 0.01, False, False, device=local, seed=10)
        m1.set_stream(stream.cuda_stream)
        m1.update_many(np.arange(30, dtype=np.int32) % 15)
        k1 = 3000
        ms1 = timed(lambda: m1.update_many(np.arange(k1, dtype=np.int32) % 15))
        also["c1_frey"] = {"workload": "c1: Frey-shape Gaussian decoder 560-200-2, M=100, L=1 (replica per GPU)",
                           "value": world * 100 * k1 / (ms1 * 1e-3), "unit": "datapoints/s", "ms_per_step": ms1 / k1}
        m1.close()

    line = None
    if rank == 0:
        peaks = measured_peaks()
        # ---- the dominant kernel: the fused step kernel IS the timed region (one launch = K updates),
        # so its duration is the CUDA-event time above.  Algorithmic bytes per update:
        # Adagrad 20 B/parameter + the minibatch 4*M*D.  Per-phase times come from %globaltimer stamps
        # taken inside the kernel at every grid barrier (CTA 0), averaged over 50 updates.
        n_params = sum(int(np.prod(sh)) for sh in model._shapes)
        bytes_per_update = 20.0 * n_params + 4.0 * M * D
        ach = bytes_per_update * K / (ms * 1e-3) / 1e9
        phases = model.profile_update(index=3, iters=50)
        tot = sum(p[1] for p in phases)
        roofline = {"kernel": "fs::fused_step_kernel (persistent cooperative kernel; one launch = %d updates)" % K,
                    "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"],
                    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel
                    # (profiles/r1_fused_step_ncu_full.txt: 24.91 MB for a 50-update launch), scaled to K updates
                    "traffic": 24.906496e6 / 50.0 * K,
                    "peak_source": peaks["source"] + " (copy bandwidth, MEASURED_PEAKS.json)",
                    "ms_per_launch": ms, "bytes_per_launch": bytes_per_update * K,
                    "bytes_per_update": bytes_per_update, "updates_per_launch": K,
                    "tflops_whole_step": value / world * FLOPS_PER_DATAPOINT / 1e12,
                    "note": "M=100 is latency bound, not roofline bound: 8 grid barriers + 0.41 GFLOP of fp32 FFMA "
                            "per update; parameters/ADA (13 MB) stay L2 resident, so DRAM traffic is far below the "
                            "algorithmic bytes.  See DESIGN.md and profiles/",
                    "phases": [{"name": p[0], "us": round(1e3 * p[1], 2), "share": round(p[1] / tot, 3),
                                "tflops": round(p[2] / (p[1] * 1e-3) / 1e12, 3) if p[2] else None,
                                "l2_gbs": round(p[3] / (p[1] * 1e-3) / 1e9, 1)} for p in phases]}
        # ---- CPU baseline on the host cores (bounded sample) -------------------------------
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            done, dt = cpu_step_loop(x[:5000], 100000, 5, budget_s=12.0)
            cpu = {"value": done * M / dt, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                   "sample": "%d update() steps (%.1f s) of the numpy fp32 restatement in oracle/" % (done, dt)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (hi+lo operands, fp32 accumulate)", "bf16": "bf16"}[args.precision],
                "data": "synthetic",
                "config": {"workload": "c2: MNIST-shape Bernoulli VAE D=784 H=500 Z=20, M=100, L=1, Adagrad; one step "
                                       "= one update() on 100 rows; N>1 = independent replicas (M=100 does not shard)",
                           "rows_resident": N_TRAIN,
                           "l2": "x_train (157 MB) exceeds the 126 MB L2 and minibatches are visited in shuffled "
                                 "order; the 3.3 MB parameter/ADA/gradient buffers stay L2-resident as in real training",
                           "eps": "Philox4x32-10 on device", "precision": args.precision},
                "clocks": clocks_summary,
                "warmup_probe_ms_per_step": probe_ms_per_step, "remeasured": remeasured,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": M * D * 4, "d2h_bytes_per_step": 4,
                        "steps": Ka, "ms_per_step": ms_e2e / Ka,
                        "api": "VAEB.update_host_async + collect -> vaeb_update_host_async / vaeb_collect (pinned host minibatch "
                               "per step, H2D on a copy stream overlapping the previous step's kernel, 4-byte D2H of every bound)",
                        "sync_call": {"value": e2e_sync, "ms_per_step": ms_e2e_sync / Ke, "steps": Ke,
                                      "api": "VAEB.update_host -> vaeb_update_host (H2D, update, D2H, host sync every step)"}},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "final_bound_per_datapoint": float(np.mean(elbos[-50:])),
                "flops_per_datapoint": FLOPS_PER_DATAPOINT,
                "achieved_tflops_whole_step": value / world * FLOPS_PER_DATAPOINT / 1e12, "also": also}
    model.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the c1/c3/c5 side measurements")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16x3", "bf16"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
