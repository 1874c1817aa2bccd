#!/usr/bin/env python
"""bench.py -- molecules/sec per GAN step (G+D fwd+bwd, WGAN-GP) at N=45, 8-layer encoder.

    python bench.py --gpus N --steps K --warmup W            # our B200 path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU cores

A "step" is one train.py:351-384 iteration on one synthetic batch: D loss (real, fake, gradient
penalty with its double backward) -> backward -> AdamW; G loss -> backward -> AdamW; two .item() syncs.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")   # B=2048 peaks at ~143 GB: avoid fragmentation
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_DIM, B_DIM, DIM, HEADS, MLP_RATIO = 13, 5, 128, 8, 3
METRIC = "molecules/sec per GAN step (G+D fwd+bwd) at N=45, 8-layer encoder"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(kernel_key, rows):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/r01b_traffic.json, taken at 1 036 800 edge rows), scaled linearly to this run's rows; None if not captured."""
    tab = None
    for name in ("r02e_traffic.json", "r02b_traffic.json", "r02_traffic.json", "r01b_traffic.json", "r01_traffic.json"):         # newest capture first
        try:
            tab = json.load(open(os.path.join(ROOT, "profiles", name)))
            break
        except Exception:
            continue
    if tab is None:
        return None
    import re
    base = re.sub(r"R=\d+,", "", kernel_key)
    base = base.replace(",bf16", "")        # variant tags (+gate, +resid, a16 ...) stay: only the plain captured kernels match
    for k, v in tab.items():
        if not k.startswith("_") and k == base:
            return v * rows / tab["_rows"]
    return None


def gan_flops_per_molecule(n, depth, d=DIM, r=MLP_RATIO, m=M_DIM, b=B_DIM):
    """Necessary GEMM FLOPs of one GAN step per molecule (74.57 GF at N=45, L=8).
    encoder fwd per layer F = d^2 [N^2 (4+4r) + N (8+4r)]; prologue P and heads counted too."""
    f_enc = d * d * (n * n * (4 + 4 * r) + n * (8 + 4 * r)) * depth
    f_d_last_dead = d * d * n * n * (2 + 4 * r)          # D's last block: out_e + mlp2 have no consumer
    pro = 2 * (n * n * (b * 64 + 64 * d) + n * (m * 64 + 64 * d))
    g_fwd = f_enc + pro + 2 * (n * d * m + n * n * d * b)
    d_fwd = f_enc - f_d_last_dead + pro + 2 * (n * d * 64 + 64 * 32 + 32 * 16 + 16)
    # D-step: D fwd x3 + G fwd; D bwd (dgrad+wgrad) x2; GP: input-bwd (1x) + double backward (~4x fwd)
    d_step = 3 * d_fwd + g_fwd + 2 * 2 * d_fwd + (1 + 4) * d_fwd
    # G-step: G fwd + D fwd + D dgrad-only bwd + G full bwd
    g_step = g_fwd + d_fwd + d_fwd + 2 * g_fwd
    return float(d_step + g_step)


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def mark(self):
        """Samples taken so far (nvidia-smi's own start-up, inside the warm-up) do not count: the record is of the timed region."""
        self.rows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        load = sm[len(sm) // 2:] if sm else []          # upper half = samples under load
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


REF_THREADS = 5          # the reference pins torch.set_num_threads(5) (train.py:16)


def cpu_gan(depth, n, sample_b, threads, skip_dead_d_grads=True):
    """The reference algorithm on host cores.  Kind "port": the reference is Python + torch_geometric / rdkit drivers and
    does not travel to the GPU box; oracle/encoder_oracle.py restates its src/model + loss.py + AdamW step with the same ATen
    ops (F.linear, F.layer_norm, softmax) and is pinned to it by the golden vectors.  tools/cpu_arm_check.py times it against
    the unmodified reference modules in the build container: 1.02x the reference's speed (profiles/r02_cpu_arm_check.json)."""
    import druggen_b200 as dg
    from oracle import encoder_oracle as orc
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    G = dg.Generator("relu", n, B_DIM, M_DIM, 0.0, dim=DIM, depth=depth, heads=HEADS, mlp_ratio=MLP_RATIO)
    D = dg.Discriminator("relu", n, B_DIM, M_DIM, 0.0, dim=DIM, depth=depth, heads=HEADS, mlp_ratio=MLP_RATIO)
    gan = orc.OracleGAN(dict(G.state_dict()), dict(D.state_dict()), depth, depth, HEADS)
    a, x = orc.synthetic_batch(sample_b, n, M_DIM, B_DIM, seed=1)

    def step():
        return gan.step(a, x, a, x, torch.rand(sample_b, 1, 1, 1), torch.rand(sample_b, 1, 1), skip_dead_d_grads=skip_dead_d_grads)
    return step


def time_cpu(step, reps, warm=1):
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    return (time.perf_counter() - t0) / reps


def cpu_baseline_record(args, reps, warm=1):
    """All host cores and, beside it, the reference's own 5-thread setting."""
    threads = os.cpu_count() or 1
    skip = not args.keep_dead_d_grads
    dt = time_cpu(cpu_gan(args.depth, args.atoms, args.cpu_sample, threads, skip), reps, warm)
    used = torch.get_num_threads()
    dt5 = time_cpu(cpu_gan(args.depth, args.atoms, args.cpu_sample, min(REF_THREADS, threads), skip), max(1, reps // 2), 1)
    torch.set_num_threads(threads)
    return dt, {"value": args.cpu_sample / dt, "unit": "molecules/s", "cores": used, "kind": "port",
                "sample": f"{reps} steps of {args.cpu_sample} molecules (same GAN step, depth {args.depth}, N={args.atoms}, fp32 PyTorch "
                          f"CPU + AdamW; the reference needs ~0.8 GB RAM per molecule at depth 8)",
                "value_at_reference_threads": args.cpu_sample / dt5, "reference_threads": min(REF_THREADS, threads),
                "port_vs_real_reference_speed": 1.02, "dead_d_wgrads_in_g_step": "computed" if args.keep_dead_d_grads else "not computed"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dt, rec = cpu_baseline_record(args, args.steps, args.warmup)
    val = rec["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "molecules/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.batch),      # the arm's config; the CPU runs a bounded sample of it (below)
        "cpu_baseline": rec,
        "e2e": {"value": val, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, batch_per_gpu):
    return {"workload": f"DrugGEN-{args.workload} GAN train step (train.py:351-384), Generator+Discriminator, "
                        f"{args.depth} encoder layers, N={args.atoms}, dim {DIM}, heads {HEADS}, mlp_ratio {MLP_RATIO}",
            "batch_per_gpu": batch_per_gpu, "atoms": args.atoms, "depth": args.depth, "precision": args.precision,
            "parallelism": f"dp{args.gpus}", "l2": "inputs_exceed_l2", "wire_format": getattr(args, "wire", "labels"),
            # train.py:371-377 also fills D's .grad in the G-step; reset_grad (train.py:352) discards it unread
            "dead_d_wgrads_in_g_step": "computed" if getattr(args, "keep_dead_d_grads", False) else "not launched",
            # GANTrainer(sequenced=True): real / fake / gradient-penalty terms backpropagated one after the other; blocks keep their
            # forward intermediates while the device has the safety margin free, else their backward recomputes (same results)
            # who issues a block's launches: the library (one C call per block and direction) or block.py, launch by launch
            "block_sequencing": ("python (one call per launch)" if os.environ.get("DRUGGEN_B200_NATIVE_BLOCK", "1") == "0" or
                                 args.precision != "bf16" else "dg_block_fwd / dg_block_bwd / dg_block_bwd_bwd"),
            "activations": "recomputed" if os.environ.get("DRUGGEN_B200_KEEP", "1") == "0" else
                           f"kept while >= {os.environ.get('DRUGGEN_B200_KEEP_HEADROOM_GB', '40')} GB free, else recomputed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=2048, help="molecules per GPU per step")
    ap.add_argument("--atoms", type=int, default=45)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--workload", default="AKT1", choices=["NoTarget", "AKT1"])
    ap.add_argument("--precision", default=os.environ.get("DRUGGEN_B200_PRECISION", "bf16"))
    ap.add_argument("--cpu-sample", type=int, default=8)
    ap.add_argument("--wire", default="labels", choices=["labels", "onehot"],
                    help="host->device format of a molecule batch: uint8 labels [B,N,N] / [B,N] (1 byte per edge) or the fp32 "
                         "one-hot tensors load_molecules builds on the host (20 bytes per edge)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-table", action="store_true", help="add the per-kernel time table of one warm-up step")
    ap.add_argument("--keep-dead-d-grads", action="store_true",
                    help="also launch the Discriminator weight gradients of the G-step that train.py computes and never reads")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import druggen_b200 as dg
    from druggen_b200 import _lib, gan, parallel

    rank, world, local = parallel.init_from_env("nccl")
    assert world == args.gpus or world == 1, (world, args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dg.set_precision(args.precision)
    be = _lib.cuda_backend()

    torch.manual_seed(0)                      # replicated weights
    n, bsz = args.atoms, args.batch
    G = dg.Generator("relu", n, B_DIM, M_DIM, 0.0, dim=DIM, depth=args.depth, heads=HEADS, mlp_ratio=MLP_RATIO).to(dev)
    D = dg.Discriminator("relu", n, B_DIM, M_DIM, 0.0, dim=DIM, depth=args.depth, heads=HEADS, mlp_ratio=MLP_RATIO).to(dev)
    trainer = gan.GANTrainer(G, D, skip_dead_d_grads=not args.keep_dead_d_grads)
    torch.manual_seed(1234 + rank)            # per-rank GP eps stream
    as_labels = args.wire == "labels"
    mol_a_h, mol_x_h = gan.synthetic_molecules(bsz, n, M_DIM, B_DIM, seed=1 + rank, labels=as_labels)
    host = [mol_a_h.pin_memory(), mol_x_h.pin_memory()]
    if args.workload == "AKT1":               # DrugGEN submodel: independent "real drug" batch (train.py:340-342)
        da, dx = gan.synthetic_molecules(bsz, n, M_DIM, B_DIM, seed=1001 + rank, labels=as_labels)
        host += [da.pin_memory(), dx.pin_memory()]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)

    def upload():
        t = [h.to(dev, non_blocking=True) for h in host]
        return (t[2], t[3], t[0], t[1]) if len(t) == 4 else (t[0], t[1], t[0], t[1])

    resident = upload()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (one of these steps also times every launch, to find the dominant kernel)
    first_step_table, dominant = {}, None
    prof_step = min(1, args.warmup - 1)       # not the very first step: its lazy initialisations (allocator growth, module
    clocks = ClockSampler(local)
    for i in range(args.warmup):              # loads) would be billed to whichever launches happen to follow them
        if rank == 0 and i == args.warmup - 1:
            clocks.start()                    # nvidia-smi starts (NVML load, ~1 s of host and driver time) under the last warm-up
        be.profile_all = i == prof_step       # step, not inside the timed region; its samples before mark() are dropped
        trainer.step(*resident)
        if i == prof_step:
            torch.cuda.synchronize()
            table = be.profile_summary()
            first_step_table = {k: {"n": v["n"], "ms": round(v["ms"], 3), "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                                    "TFLOPs": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1)} for k, v in table.items()}
            be.profile_all = False
            dominant = max(table, key=lambda k: table[k]["ms"]) if table else None
            be.profile_only = dominant
    barrier()
    be.profile_reset()                        # the dominant kernel's events below are those of the TIMED steps only

    # ---- timed: device-resident inputs
    if rank == 0:
        if clocks.proc is None:
            clocks.start()
        clocks.mark()
    launches0 = be.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        losses = trainer.step(*resident)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = (be.launches - launches0) // args.steps
    dom = be.profile_summary().get(be.profile_only) if be.profile_only else None      # events of the TIMED steps only
    be.profile_only = None
    dg.kernels.check_labels()

    # ---- timed: end to end through the public API with host (pinned) inputs
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(args.steps):
        losses = trainer.step(*upload())
    ev3.record()
    barrier()
    ms_e2e = ev2.elapsed_time(ev3) / args.steps
    clk = clocks.stop() if rank == 0 else None

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank != 0:
        torch.distributed.destroy_process_group()
        return
    total = bsz * world
    pk, pk_kind = peaks()
    flops_mol = gan_flops_per_molecule(n, args.depth)
    out = {
        "metric": METRIC, "value": total / (ms / 1e3), "unit": "molecules/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "fp32": "f32", "bf16x3": "bf16x3"}[args.precision], "data": "synthetic",
        "config": workload_config(args, bsz), "clocks": clk,
        "e2e": {"value": total / (ms_e2e / 1e3), "unit": "molecules/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8},
        "gpu_launches": launches, "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
        "reserved_mem_gb": round(torch.cuda.max_memory_reserved() / 2 ** 30, 1),
        "allocator": {k: torch.cuda.memory_stats().get(v, 0) for k, v in (("alloc_retries", "num_alloc_retries"), ("ooms", "num_ooms"),
                                                                          ("device_mallocs", "num_device_alloc"), ("device_frees", "num_device_free"))},
        "losses": {"d": losses[0], "g": losses[1]},
    }
    if args.kernel_table:
        out["kernel_table"] = dict(sorted(first_step_table.items(), key=lambda kv: -kv[1]["ms"]))
    step_tflops = flops_mol * bsz / (ms / 1e3) / 1e12
    if dom:
        # the dominant kernel against BOTH roofs: tensor pipe on its algorithmic FLOPs, HBM on its algorithmic
        # bytes (the kernel's inputs + outputs; operand spills such as h / dh do not count) and on the bytes it really moves
        sec = dom["ms"] / 1e3
        tf, gb_alg, gb_impl = dom["flops"] / sec / 1e12, dom["alg_bytes"] / sec / 1e9, dom["bytes"] / sec / 1e9
        t_frac, h_frac = tf / pk["bf16_tflops_sustained"], gb_alg / pk["hbm_gbs"]
        bound = "tensor" if (dom["flops"] > 0 and t_frac >= h_frac) else "hbm"       # the roof that allows the least
        ach, peak, unit = (tf, pk["bf16_tflops_sustained"], "TFLOP/s") if bound == "tensor" else (gb_alg, pk["hbm_gbs"], "GB/s")
        out["roofline"] = {"kernel": be.profile_name(dom), "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                           "frac": ach / peak, "traffic": ncu_traffic(be.profile_name(dom), bsz * n * n), "peak_source": pk_kind,
                           "tensor_frac": t_frac, "tensor_tflops": tf, "hbm_frac_algorithmic": h_frac,
                           "hbm_frac_implemented": gb_impl / pk["hbm_gbs"], "hbm_gbs_implemented": gb_impl,
                           "launches": dom["n"], "avg_launch_ms": dom["ms"] / dom["n"], "share_of_step": dom["ms"] / (ms * args.steps),
                           "algorithmic_bytes_per_launch": dom["alg_bytes"] / dom["n"],
                           "implemented_bytes_per_launch": dom["bytes"] / dom["n"],
                           "algorithmic_flops_per_launch": dom["flops"] / dom["n"],
                           "step_tflops": step_tflops, "step_frac_of_bf16_sustained": step_tflops / pk["bf16_tflops_sustained"],
                           "step_flops_per_molecule": flops_mol}
    if os.environ.get("DRUGGEN_BENCH_RETRY") == "1":
        out["memory_fallback"] = {"kept_intermediates": False, "why": "CUDA OOM with the blocks keeping their forward intermediates; "
                                  "same batch, recomputing backward (block.keep_intermediates off)"}
    elif os.environ.get("DRUGGEN_BENCH_RETRY"):
        out["batch_fallback"] = {"requested": 2048, "ran": bsz, "why": "CUDA OOM at the requested batch on this device"}
    if world == 1 and not args.no_cpu_baseline:
        _, out["cpu_baseline"] = cpu_baseline_record(args, 2)
    print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except torch.cuda.OutOfMemoryError:
        # the default run (2048 molecules per GPU; the blocks keep forward intermediates while >= 40 GB stay free, ~150 GB peak)
        # did not fit next to whatever else holds memory on this device: re-run in a fresh process, first with the recomputing
        # backward at the same batch (~71 GB peak), then at half the batch.  The JSON line says which ran (memory_fallback /
        # batch_fallback keys).
        level = os.environ.get("DRUGGEN_BENCH_RETRY")
        if "--batch" in sys.argv or level == "2" or int(os.environ.get("WORLD_SIZE", "1")) > 1:
            raise
        if level is None:
            os.environ["DRUGGEN_BENCH_RETRY"], os.environ["DRUGGEN_B200_KEEP"] = "1", "0"
            os.execv(sys.executable, [sys.executable] + sys.argv)
        os.environ["DRUGGEN_BENCH_RETRY"] = "2"
        os.execv(sys.executable, [sys.executable] + sys.argv + ["--batch", "1024"])
