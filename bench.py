#!/usr/bin/env python
"""bench.py -- headline benchmark of the q100 hot path (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the seven QuantLinear GEMMs of ONE LLaMA3-8B decoder
block, W4A16 g128, random-init weights, synthetic activations.  One "step" = one prefill
pass (M = 2048 tokens) over q/k/v/o/gate/up/down.  Metric: dequant-GEMM TFLOP/s
(2*M*N*K summed over the seven linears / step time).  The same JSON line also carries the
decode-shape numbers (GB/s of packed-weight stream, M = 1 and 16) and the quantize kernel's
GB/s, each with its roofline fraction.

N > 1 (torchrun): tensor parallel -- q/k/v/gate/up column-sharded, o/down row-sharded, one
NCCL all-reduce after each row-parallel GEMM.  The token batch grows with the GPU count
(M = 2048 * N), so every GPU keeps the single-GPU amount of GEMM work: scaling is "weak".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "llama3-quantization_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

HIDDEN, INTER, KV = 4096, 14336, 1024
# name -> (N, K, input buffer)
LINEARS = [("q_proj", HIDDEN, HIDDEN, "h"), ("k_proj", KV, HIDDEN, "h"), ("v_proj", KV, HIDDEN, "h"),
           ("o_proj", HIDDEN, HIDDEN, "a"), ("gate_proj", INTER, HIDDEN, "h"), ("up_proj", INTER, HIDDEN, "h"),
           ("down_proj", HIDDEN, INTER, "m")]
PREFILL_M = 2048
GROUP, WBITS = 128, 4
METRIC = "w4a16_g128_dequant_gemm_tflops"
UNIT = "TFLOP/s"


def WORKLOAD_CONFIG(m_total, world):
    """The same `config` object on both arms (the driver compares them): BASELINE.json configs[1]."""
    return {"workload": "llama3-8b decoder block: 7 QuantLinear GEMMs (q,k,v,o,gate,up,down), W4A16 g128, prefill, "
                        "2048 tokens per GPU",
            "tokens": m_total, "parallelism": f"tp{world}" if world > 1 else "single"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tc_burst": d["bf16_tflops"], "tc_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "tc_burst": 1590.0, "tc_sustained": 1400.0, "src": "fallback"}


def ncu_traffic(key):
    """dram bytes per launch of a kernel from the committed ncu --set full digest (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return float(json.load(open(p))[key]["dram_bytes_per_launch"])
    except Exception:                                               # noqa: BLE001
        return None


def step_flops(M=PREFILL_M):
    return sum(2.0 * M * n * k for _, n, k, _ in LINEARS)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path, on the host cores.

    oracle/_ref holds the reference's unmodified modules (vendored by oracle/build_ref.py; they travel to the GPU
    box with the snapshot): this leg builds the reference's QuantLinear (quant/int_linear.py) around the seven
    linears of the block, fake-quantises the weights in place with its UniformAffineQuantizer exactly like
    quant/utils.py::smooth_and_quant_inplace does after calibration, and times QuantLinear.forward
    (int_linear.py:48-65: F.linear on the dense fake-quantised weight) at M = 2048 with every host thread.
    fp32 on the CPU (the reference keeps fp16 for the GPU only).  Falls back to oracle/torch_port.py (kind "port")
    only when oracle/_ref is missing."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    gen = torch.Generator().manual_seed(0)
    M = PREFILL_M
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    kind = "reference" if os.path.exists(os.path.join(ref_dir, "quant", "int_linear.py")) else "port"
    wparams = {"n_bits": WBITS, "per_channel_axes": [0], "symmetric": False, "dynamic_method": "per_channel",
               "group_size": GROUP, "lwc": False, "disable_zero_point": False}
    aparams = {"n_bits": 16, "per_channel_axes": [], "symmetric": False, "dynamic_method": "per_token"}
    with torch.no_grad():
        if kind == "reference":
            pkg = os.path.join(ROOT, "llama3-quantization_b200")
            sys.path[:] = [p for p in sys.path if os.path.abspath(p) != pkg]      # never the drop-in on this arm
            sys.path[:0] = [ref_dir, os.path.join(ref_dir, "stubs")]
            import quant.int_linear as RIL
            assert os.path.realpath(RIL.__file__).startswith(os.path.realpath(ref_dir)), RIL.__file__
            mods = {}
            for j, (name, n, k, _) in enumerate(LINEARS):
                lin = torch.nn.Linear(k, n, bias=False)
                lin.weight.copy_(torch.randn(n, k, generator=gen) * 0.02)
                ql = RIL.QuantLinear(lin, wparams, aparams)
                ql.weight = ql.weight_quantizer(ql.weight)             # quant/utils.py:133-136
                ql.set_quant_state(weight_quant=False, act_quant=False)
                mods[name] = ql
            xs = {"h": torch.randn(M, HIDDEN, generator=gen), "a": torch.randn(M, HIDDEN, generator=gen),
                  "m": torch.randn(M, INTER, generator=gen)}

            def step():
                for name, n, k, src in LINEARS:
                    mods[name](xs[src])
        else:
            import torch_port as TP
            weights = {}
            for j, (name, n, k, _) in enumerate(LINEARS):
                w = torch.randn(n, k, generator=gen) * 0.02
                weights[name] = TP.fake_quant(w, WBITS, GROUP)[0]          # post-omniquant state
            xs = {"h": torch.randn(M, HIDDEN, generator=gen), "a": torch.randn(M, HIDDEN, generator=gen),
                  "m": torch.randn(M, INTER, generator=gen)}

            def step():
                for name, n, k, src in LINEARS:
                    TP.quant_linear(xs[src], weights[name], None, WBITS, GROUP, requantize=False)

        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
    val = step_flops() / dt / 1e12
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": WORKLOAD_CONFIG(PREFILL_M * max(args.gpus, 1), max(args.gpus, 1)),
            "notes": {"where": "host CPU, the reference's own QuantLinear.forward (torch eager, fp32)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                             "sample": f"7 linears at M=2048 of the {PREFILL_M * max(args.gpus, 1)} tokens per step "
                                       "(the CPU rate does not depend on M at this size), fp32, torch CPU"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ our arm
def timed(fn, iters, stream=None):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / iters          # ms


def run_ours(args):
    import torch.distributed as dist
    import q100
    from q100 import tp, _lib
    from quant.int_linear import QuantLinear

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the q100 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        if not os.environ.get("Q100_NO_PEER"):
            # K7: the small all-reduces of the decode steps go over peer memory instead of NCCL (falls back by itself)
            from q100 import peer
            peer.enable(None, dev)
            # K7b: at prefill sizes the reduce-scatter is the row-parallel GEMM's epilogue (q100/peer.py ScatterReduce)
            if not os.environ.get("Q100_NO_SCATTER"):
                peer.enable_scatter(None, dev, PREFILL_M * world, HIDDEN, torch.bfloat16)
                peer.enable_scatter(None, dev, PREFILL_M, 8192, torch.float16)       # the model rows (fp16, seq 2048)
    P = peaks()
    dt_t = torch.bfloat16
    flags = _lib.GEMM_FORCE_SKINNY if os.environ.get("Q100_BENCH_FORCE_SKINNY") else 0

    # ---- weights: NB independent blocks so that successive decode steps stream from HBM, not L2.
    # Besides the seven per-linear packs (prefill / module API) each block carries the decode-time fusion
    # q|k|v and gate|up (concatenation along N is exact: groups run along K), SURVEY.md section 8.
    NB = 3
    FUSED = [("qkv", ("q_proj", "k_proj", "v_proj"), "h"), ("o", ("o_proj",), "a"),
             ("gateup", ("gate_proj", "up_proj"), "h"), ("down", ("down_proj",), "m")]
    gen = torch.Generator(device=dev).manual_seed(0)
    blocks, fblocks = [], []
    for b in range(NB):
        blk, shards = {}, {}
        for j, (name, n, k, _) in enumerate(LINEARS):
            mode = tp.LLAMA_PLAN[name]
            sh = tp.shard_bounds(n, k, mode, rank, world, GROUP)
            gen.manual_seed(b * 7 + j)                              # same full weight on every rank
            w = (torch.randn(n, k, generator=gen, device=dev) * 0.02).to(dt_t)
            wl = w[sh.n0:sh.n1, sh.k0:sh.k1].contiguous()
            pw, _ = q100.pack_weight(wl, WBITS, GROUP)
            blk[name] = (pw, mode)
            shards[name] = wl
            del w
        blocks.append(blk)
        fblk = {}
        for fname, parts, _ in FUSED:
            wl = torch.cat([shards[p_] for p_ in parts], dim=0) if len(parts) > 1 else shards[parts[0]]
            fblk[fname] = (q100.pack_weight(wl, WBITS, GROUP)[0], tp.LLAMA_PLAN[parts[0]])
        fblocks.append(fblk)
        del shards
    torch.cuda.synchronize()

    def make_inputs(M):
        g2 = torch.Generator(device=dev).manual_seed(1)
        full = {"h": torch.randn(M, HIDDEN, generator=g2, device=dev).to(dt_t),
                "a": torch.randn(M, HIDDEN, generator=g2, device=dev).to(dt_t),
                "m": torch.randn(M, INTER, generator=g2, device=dev).to(dt_t)}
        return {k: v for k, v in full.items()}

    launches = [0]
    scatter = None
    if world > 1:
        from q100 import peer as _peer0
        scatter = _peer0.get_scatter(None, dt_t)

    def block_step(blk, xs):
        outs = []
        for name, n, k, src in LINEARS:
            pw, mode = blk[name]
            x = tp.shard_input(xs[src], mode, rank, world) if world > 1 else xs[src]
            sc = scatter if (world > 1 and mode == tp.ROW) else None
            if sc is not None and sc.fits(x.shape[0], pw):
                y = sc(x, pw)                       # GEMM with the reduce-scatter as its epilogue + reduce / broadcast
                launches[0] += 2
            else:
                y = q100.gemm_wna16(x, pw, None, flags)
                launches[0] += 1
                tp.reduce_output(y, mode)
            outs.append(y)
        return outs

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # =============== headline: prefill, M = 2048 tokens per GPU ===============
    M_PRE = PREFILL_M * world
    xs = make_inputs(M_PRE)
    it = [0]

    def prefill_step():
        block_step(blocks[it[0] % NB], xs)
        it[0] += 1

    for _ in range(max(args.warmup, 3)):
        prefill_step()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    launches[0] = 0
    ms = timed(prefill_step, args.steps)
    sync_all()
    clocks = sampler.stop() if sampler else None
    n_launch = launches[0]
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = step_flops(M_PRE) / (ms * 1e-3) / 1e12

    # per-launch duration of the dominant kernel (tcgen05 dequant-GEMM on up_proj), events on the launch stream
    pw_up, mode_up = blocks[0]["up_proj"]
    x_up = tp.shard_input(xs["h"], mode_up, rank, world) if world > 1 else xs["h"]
    ms_up = min(timed(lambda: q100.gemm_wna16(x_up, pw_up, None, flags), 10) for _ in range(3))
    fl_up = 2.0 * M_PRE * pw_up.N * pw_up.K
    ach = fl_up / (ms_up * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "k3_tca (W4A16 g128 dequant-GEMM, weights as the TMEM A operand, "
                                             f"up_proj {pw_up.N}x{pw_up.K} per GPU, M={M_PRE})",
                "achieved": ach, "peak": P["tc_burst"], "unit": "TFLOP/s", "frac": ach / P["tc_burst"],
                "peak_src": P["src"] + " cuBLAS bf16 burst",
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel (single GPU, up_proj,
                # M = 2048) from the committed ncu --set full capture (profiles/ncu_traffic.json names the capture and
                # the command); null for shapes the capture does not cover
                "traffic": ncu_traffic("k3_tca_up_proj_m2048") if (world == 1 and M_PRE == 2048) else None,
                "traffic_unit": "bytes per launch", "traffic_src": "profiles/ncu_traffic.json",
                "frac_of_sustained": ach / P["tc_sustained"]}

    # =============== decode shapes: packed-weight stream GB/s ===============
    # One decode step of the block's linears = 4 launches (q|k|v, o, gate|up, down), captured for all NB
    # weight sets in one CUDA graph with programmatic dependent launch, so that each kernel's weight
    # stream starts while its predecessor drains.  "api" = the same step as 7 separate un-graphed calls.
    decode = {}
    dflags = flags | _lib.GEMM_PDL
    for M in (1, 4, 16):
        xd = make_inputs(M)

        def fused_step(fb):
            for fname, parts, src in FUSED:
                pw, mode = fb[fname]
                x = tp.shard_input(xd[src], mode, rank, world) if world > 1 else xd[src]
                y = q100.gemm_wna16(x, pw, None, dflags)
                tp.reduce_output(y, mode)

        for fb in fblocks:
            fused_step(fb)
        sync_all()
        graph = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fused_step(fblocks[0])
            torch.cuda.current_stream().wait_stream(side)
            sync_all()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for fb in fblocks:
                    fused_step(fb)
            graph.replay()
            sync_all()
            msd = timed(graph.replay, 20) / NB
        except Exception as ex:                                   # e.g. a collective that cannot be captured
            graph = None
            jt = [0]

            def dstep():
                fused_step(fblocks[jt[0] % NB])
                jt[0] += 1
            msd = timed(dstep, 30)
        t = torch.tensor([msd], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        msd = float(t.item())
        esz = 2
        byts = sum(fblocks[0][fn][0].nbytes() + M * fblocks[0][fn][0].K * esz + M * fblocks[0][fn][0].N * esz
                   for fn, *_ in FUSED)
        gbs = byts * world / (msd * 1e-3) / 1e9
        # per-launch time of the largest fused kernel (gate|up: 3 x 62 MB rotate through more than the L2),
        # 12 launches in one graph
        per = {}
        if rank == 0:
            fname, src = "gateup", "h"
            x = tp.shard_input(xd[src], fblocks[0][fname][1], rank, world) if world > 1 else xd[src]
            for i in range(3):
                q100.gemm_wna16(x, fblocks[i % NB][fname][0], None, dflags)
            torch.cuda.synchronize()
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                for i in range(12):
                    q100.gemm_wna16(x, fblocks[i % NB][fname][0], None, dflags)
            g1.replay()
            torch.cuda.synchronize()
            m1 = timed(g1.replay, 10) / 12
            nb_ = fblocks[0][fname][0].nbytes()
            per[fname] = {"us": m1 * 1e3, "GBps": nb_ / (m1 * 1e-3) / 1e9, "frac_hbm": nb_ / (m1 * 1e-3) / 1e9 / P["hbm"],
                          "MB": nb_ / 1e6}
            del g1
        jt = [0]

        def api_step():
            block_step(blocks[jt[0] % NB], xd)
            jt[0] += 1
        for _ in range(3):
            api_step()
        sync_all()
        ms_api = timed(api_step, 20)
        decode[f"M{M}"] = {"ms_per_block": msd, "algorithmic_MB_per_gpu": byts / 1e6, "GBps": gbs,
                           "frac_hbm": gbs / (P["hbm"] * world), "launches_per_block": len(FUSED),
                           "cuda_graph": graph is not None, "pdl": True,
                           "linear_tokens_per_s_32_blocks": M / (32 * msd * 1e-3),
                           "per_launch": per, "api_7_launches_ms_per_block": ms_api}
        del graph

    # =============== quantize kernel (K1): fake-quant and pack modes on up_proj ===============
    quant = {}
    if rank == 0:
        wq = [(torch.randn(INTER, HIDDEN, device=dev) * 0.02).to(torch.float16) for _ in range(3)]
        qi = [0]

        # inputs AND outputs rotate through three buffers each (3 x 117 MB in, 3 x 117 / 29 MB out): the last three
        # results stay alive, so the allocator cannot hand the same (L2-resident) block back and the write stream
        # has to reach HBM like the read stream does
        keep = [None, None, None]

        def q_fake():
            keep[qi[0] % 3] = q100.quantize_groupwise(wq[qi[0] % 3], GROUP, WBITS, want_dequant=True)
            qi[0] += 1

        def q_pack():
            keep[qi[0] % 3] = q100.quantize_groupwise(wq[qi[0] % 3], GROUP, WBITS, want_dequant=False, want_packed=True)
            qi[0] += 1

        n_el = INTER * HIDDEN
        for nm, fn, bpe in (("fake_quant_f16", q_fake, 4.03125), ("pack_f16", q_pack, 2.53125)):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            msq = timed(fn, 30)
            gbs = n_el * bpe / (msq * 1e-3) / 1e9
            quant[nm] = {"ms": msq, "GBps": gbs, "frac_hbm": gbs / P["hbm"], "bytes_per_element": bpe,
                         "outputs": "rotated over 3 live buffers"}
            keep[:] = [None, None, None]
        del wq
        # K2: per-token activation fake-quant (the act_quantizer of the W8A8 / W4A4 configurations), prefill-size input
        xa = [torch.randn(M_PRE, HIDDEN, device=dev).to(torch.float16) for _ in range(3)]

        def q_act():
            q100.fakequant_rows(xa[qi[0] % 3], 8)
            qi[0] += 1

        for _ in range(5):
            q_act()
        torch.cuda.synchronize()
        msq = timed(q_act, 30)
        bpe = 4.0 + 4.0 / HIDDEN
        gbs = M_PRE * HIDDEN * bpe / (msq * 1e-3) / 1e9
        quant["act_fake_quant_f16_a8"] = {"ms": msq, "GBps": gbs, "frac_hbm": gbs / P["hbm"], "bytes_per_element": bpe,
                                          "shape": [M_PRE, HIDDEN],
                                          "note": "through the Python API: host-bound at this size (allocations + ctypes)"}
        # the same 12 calls captured in a CUDA graph: the kernel without the per-call host work
        try:
            sg = torch.cuda.Stream()
            sg.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(sg):
                q_act()
            torch.cuda.current_stream().wait_stream(sg)
            torch.cuda.synchronize()
            gq = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gq):
                for _ in range(12):
                    q_act()
            gq.replay()
            torch.cuda.synchronize()
            msg_ = timed(gq.replay, 5) / 12
            gbs_g = M_PRE * HIDDEN * bpe / (msg_ * 1e-3) / 1e9
            quant["act_fake_quant_f16_a8_graph"] = {"ms": msg_, "GBps": gbs_g, "frac_hbm": gbs_g / P["hbm"],
                                                    "bytes_per_element": bpe, "shape": [M_PRE, HIDDEN]}
            del gq
        except Exception as ex:                                     # noqa: BLE001
            quant["act_fake_quant_f16_a8_graph"] = {"error": str(ex)[:200]}
        del xa
        # K1b: fused backward of the quantizer (calibration loop), fp32 like the reference's training dtype, LWC on
        from q100 import ops as _qops
        xb = torch.randn(INTER, HIDDEN, device=dev) * 0.02
        gb = torch.randn(INTER, HIDDEN, device=dev)
        ngr = INTER * HIDDEN // GROUP
        sub_, slb_ = torch.rand(ngr, 1, device=dev) * 0.4 + 0.6, torch.rand(ngr, 1, device=dev) * 0.4 + 0.6

        def q_bwd():
            _qops.quantize_backward(xb, gb, GROUP, WBITS, False, False, sub_, slb_, True, True)

        for _ in range(5):
            q_bwd()
        torch.cuda.synchronize()
        msq = timed(q_bwd, 20)
        bpe = 12.0 + 16.0 / GROUP                       # x + grad_out read, grad_x written, 4 group scalars
        gbs = n_el * bpe / (msq * 1e-3) / 1e9
        quant["backward_f32_lwc"] = {"ms": msq, "GBps": gbs, "frac_hbm": gbs / P["hbm"], "bytes_per_element": bpe}
        del xb, gb, sub_, slb_

    # =============== e2e: module API, host buffers, H2D + D2H inside the timed region ===============
    e2e = None
    cpu_base = None
    if True:
        # every rank drives its shard of the block through QuantLinear.forward: H2D of the (sharded)
        # activations from pinned host memory, the GEMMs, the row-parallel all-reduce, D2H of the outputs
        mods = []
        g3 = torch.Generator(device=dev)
        for j, (name, n, k, src) in enumerate(LINEARS):
            mode = tp.LLAMA_PLAN[name]
            sh = tp.shard_bounds(n, k, mode, rank, world, GROUP)
            g3.manual_seed(j)
            wfull = (torch.randn(n, k, generator=g3, device=dev) * 0.02).to(dt_t)
            lin = torch.nn.Linear(sh.k1 - sh.k0, sh.n1 - sh.n0, bias=False, device=dev, dtype=dt_t)
            with torch.no_grad():
                lin.weight.copy_(wfull[sh.n0:sh.n1, sh.k0:sh.k1])
            del wfull
            ql = QuantLinear(lin, {"n_bits": WBITS, "per_channel_axes": [0], "symmetric": False,
                                   "dynamic_method": "per_channel", "group_size": GROUP, "lwc": False,
                                   "disable_zero_point": False},
                             {"n_bits": 16, "per_channel_axes": [], "symmetric": False, "dynamic_method": "per_token"})
            ql.set_quant_state(True, False)
            mods.append((ql, src, mode))
        # host buffers hold what THIS rank is responsible for: its K-slice of the row-parallel inputs ("a", "m"), its
        # 1/N of the token rows of the replicated input "h" (all-gathered over NVLink inside the step), its column
        # shard of q/k/v/gate/up and its 1/N of the token rows of the all-reduced o / down outputs -- every byte of the
        # step's inputs and outputs crosses PCIe exactly once over the whole job
        rows_per = M_PRE // world
        r0_, r1_ = rank * rows_per, (rank + 1) * rows_per

        def host_in(k, v):
            if world == 1:
                return v
            return v[r0_:r1_] if k == "h" else tp.shard_input(v, tp.ROW, rank, world)
        hin = {k: host_in(k, v).contiguous().cpu().pin_memory() for k, v in xs.items()}
        hout = [torch.empty((rows_per if (mode == tp.ROW and world > 1) else M_PRE), ql.out_features, dtype=dt_t).pin_memory()
                for ql, _, mode in mods]
        din = {k: torch.empty(v.shape, dtype=dt_t, device=dev) for k, v in hin.items()}
        h_full = torch.empty(M_PRE, HIDDEN, dtype=dt_t, device=dev) if world > 1 else None

        # copies ride their own streams: PCIe is full duplex and independent of the SMs, so the H2D of a later
        # input and the D2H of an earlier output overlap the GEMMs in between (still inside the timed step)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def e2e_step():
            cur = torch.cuda.current_stream()
            with torch.no_grad():
                ready = {}
                with torch.cuda.stream(s_in):
                    for k in din:
                        din[k].copy_(hin[k], non_blocking=True)
                        ready[k] = s_in.record_event()
                keep = []
                gathered = False
                for (ql, src, mode), ho in zip(mods, hout):
                    cur.wait_event(ready[src])
                    xin = din[src]
                    if src == "h" and world > 1:
                        if not gathered:
                            dist.all_gather_into_tensor(h_full, din["h"])
                            gathered = True
                        xin = h_full
                    y = tp.reduce_output(ql(xin), mode)
                    if mode == tp.ROW and world > 1:
                        y = y[r0_:r1_]
                    done = cur.record_event()
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(done)
                        ho.copy_(y, non_blocking=True)
                    y.record_stream(s_out)
                    keep.append(y)
            torch.cuda.synchronize()

        for _ in range(3):
            e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        sync_all()
        dte = (time.perf_counter() - t0) / args.steps
        te = torch.tensor([dte], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dte = float(te.item())
        e2e = {"value": step_flops(M_PRE) / dte / 1e12, "unit": UNIT, "ms_per_step": dte * 1e3,
               "h2d_bytes_per_step": sum(v.numel() * 2 for v in hin.values()) * world,
               "d2h_bytes_per_step": sum(h.numel() * 2 for h in hout) * world,
               "api": "quant.int_linear.QuantLinear.forward (use_weight_quant=True), pinned host buffers, "
                      "all ranks; bytes summed over ranks (each rank moves its own 1/N of the replicated tensors)"}
        del mods, hin, hout, din, h_full

    if rank == 0 and world == 1:
        # ---- CPU baseline on the host cores: bounded sample of the same workload (torch port)
        import torch_port as TP
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            gc = torch.Generator().manual_seed(0)
            ws = {}
            for name, n, k, _ in LINEARS:
                ws[name] = TP.fake_quant(torch.randn(n, k, generator=gc) * 0.02, WBITS, GROUP)[0]
            Ms = 512
            xc = {"h": torch.randn(Ms, HIDDEN), "a": torch.randn(Ms, HIDDEN), "m": torch.randn(Ms, INTER)}

            def cstep():
                for name, n, k, src in LINEARS:
                    TP.quant_linear(xc[src], ws[name], None, WBITS, GROUP, requantize=False)
            cstep()
            t0 = time.perf_counter()
            reps = 0
            while reps < 3 or (time.perf_counter() - t0 < 8 and reps < 50):
                cstep()
                reps += 1
            dtc = (time.perf_counter() - t0) / reps
            wq_cpu = torch.randn(INTER, HIDDEN) * 0.02
            t1 = time.perf_counter()
            TP.fake_quant(wq_cpu, WBITS, GROUP)
            tq = time.perf_counter() - t1
        cpu_base = {"value": step_flops(Ms) / dtc / 1e12, "unit": UNIT, "cores": torch.get_num_threads(),
                    "kind": "port",
                    "sample": f"7 linears of the block at M={Ms} (of 2048), fp32, {reps} reps, torch eager on host cores",
                    "fake_quant_up_proj_ms": tq * 1e3}

    # =============== the reference's deployment mode on this GPU: eager torch ops (SURVEY 8d "timing the reference") ====
    gpu_eager = None
    if rank == 0 and world == 1:
        def ref_ops_fake_quant(w):
            # the reference's op sequence written out with torch ops (quant/quantizer.py:132-159 then :94-115,
            # asymmetric, group-wise): what its UniformAffineQuantizer launches, one kernel per line
            xg = w.reshape(-1, GROUP)
            xmin, xmax = xg.amin(-1, keepdim=True), xg.amax(-1, keepdim=True)
            scale = ((xmax - xmin) / (2 ** WBITS - 1)).clamp(min=1e-5, max=1e4)
            zp = (-xmin / scale).clamp(min=-1e4, max=1e4).round()
            v = xg / scale
            x_int = ((v.round() - v).detach() + v).add(zp).clamp(0, 2 ** WBITS - 1)
            return x_int.sub(zp).mul(scale).reshape(w.shape)

        with torch.no_grad():
            wg = (torch.randn(INTER, HIDDEN, device=dev) * 0.02).half()

            def eager_fq():
                ref_ops_fake_quant(wg)
            for _ in range(2):
                eager_fq()
            ms_fq = timed(eager_fq, 5)
            wd = ref_ops_fake_quant(wg)                                  # dequantised fp16 weight, what F.linear reads
            x2k = torch.randn(PREFILL_M, HIDDEN, device=dev).half()
            x1 = torch.randn(1, HIDDEN, device=dev).half()
            wds = [wd, wd.clone(), wd.clone()]
            k_ = [0]

            def eager_lin(x):
                torch.nn.functional.linear(x, wds[k_[0] % 3])
                k_[0] += 1
            for _ in range(3):
                eager_lin(x2k); eager_lin(x1)
            ms_pre = timed(lambda: eager_lin(x2k), 10)
            ms_dec = timed(lambda: eager_lin(x1), 30)
        gpu_eager = {"what": "the reference's op sequence as eager torch ops on this GPU, up_proj 14336x4096 fp16",
                     "fake_quant_ms": ms_fq, "fake_quant_GBps_algorithmic": INTER * HIDDEN * 4.03125 / (ms_fq * 1e-3) / 1e9,
                     "prefill_linear_ms": ms_pre, "prefill_linear_tflops": 2.0 * PREFILL_M * INTER * HIDDEN / (ms_pre * 1e-3) / 1e12,
                     "decode_linear_us": ms_dec * 1e3,
                     "decode_linear_fp16_weight_GBps": INTER * HIDDEN * 2 / (ms_dec * 1e-3) / 1e9}
        del wg, wd, wds, x2k, x1

    # =============== strong scaling: the same block step at a FIXED total of 2048 tokens ===============
    strong = None
    if world > 1:
        xs_s = make_inputs(PREFILL_M)
        it_s = [0]

        def strong_step():
            block_step(blocks[it_s[0] % NB], xs_s)
            it_s[0] += 1
        for _ in range(3):
            strong_step()
        sync_all()
        ms_s = timed(strong_step, args.steps)
        sync_all()
        ts_ = torch.tensor([ms_s], device=dev, dtype=torch.float64)
        dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
        ms_s = float(ts_.item())
        strong = {"tokens": PREFILL_M, "ms_per_step": ms_s, "value": step_flops(PREFILL_M) / (ms_s * 1e-3) / 1e12,
                  "unit": UNIT, "scaling": "strong"}
        del xs_s

    # =============== whole model: random-init LLaMA3-8B W4A16 g128, tensor parallel ===========
    model_res = model70_res = None
    cfg4_res = None
    extra = {"llama3_8b": None, "llama3_8b_smoothquant_configs": None, "llama3_70b": None}

    def make_line(note=None):
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": WORKLOAD_CONFIG(M_PRE, world),
                "notes": {"l2": "3 weight sets rotated; per-step working set (109 MB packed weights + 270 MB "
                                "activations/outputs) exceeds the 126 MB L2", "force_skinny": bool(flags),
                          "row_parallel_reduce": (("through the switch: multimem.ld_reduce + multimem.st on the GEMM output "
                                                   "(K7c)" if type(scatter).__name__ == "MulticastReduce" else
                                                   "fused: scatter epilogue + reduce/broadcast over peer memory (K7b)")
                                                  if scatter is not None else "NCCL all-reduce") if world > 1 else None},
                "roofline": roofline, "decode": decode, "quantize": quant, "strong_scaling": strong,
                "llama3_8b": extra["llama3_8b"], "llama3_8b_smoothquant_configs": extra["llama3_8b_smoothquant_configs"],
                "llama3_70b": extra["llama3_70b"],
                "cpu_baseline": cpu_base, "gpu_eager_reference_ops": gpu_eager, "e2e": e2e,
                "gpu_launches": n_launch, "clocks": clocks, "peaks": P}
        if note:
            line["notes"]["model_rows"] = note
        return line

    # the whole-model rows are extras of the line: if one of them gets stuck (a lost rank, a collective that never
    # completes) the contract line still goes out with the rows that finished, and every rank leaves
    import threading
    budget_s = float(os.environ.get("Q100_BENCH_MODEL_BUDGET_S", "900"))
    done_flag = threading.Event()

    def _deadline():
        if done_flag.is_set():
            return
        if rank == 0:
            print(json.dumps(make_line(f"cut off after {budget_s:.0f} s: unfinished rows are null")), flush=True)
        os._exit(0)
    wd = threading.Timer(budget_s if rank == 0 else budget_s + 10.0, _deadline)
    wd.daemon = True
    wd.start()
    if not os.environ.get("Q100_BENCH_SKIP_MODEL"):
        blocks.clear()
        fblocks.clear()
        torch.cuda.empty_cache()
        import bench_model
        def row(**kw):
            try:
                return bench_model.run(bench_model.default_args(**kw), dev, rank, world)
            except Exception as ex:                                   # noqa: BLE001 -- the row is reported as failed, the line survives
                import traceback
                traceback.print_exc(file=sys.stderr)
                return {"error": f"{type(ex).__name__}: {ex}"[:400]}
        model_res = row(ppl_samples=1, ppl_ref=True, quant_head=not os.environ.get("Q100_BENCH_SKIP_QUANT_HEAD"))
        extra["llama3_8b"] = model_res
        # SURVEY 8d config 5 (BASELINE.json configs[4]): LLaMA3-70B W4A16 g128, tensor parallel over all ranks (the
        # packed model is 38 GB: it also fits one B200, so the row exists at every N).  Perplexity is compared on the
        # first 4 blocks (the fp16 reference of all 80 does not fit next to the model).
        # SURVEY 8d config 4 (BASELINE.json configs[3]): SmoothQuant-style W8A8 and W4A4 -- per-channel weights, per-token
        # activations (K2 in front of the prefill GEMMs and inside QuantMatMul; K3c, the integer GEMM with the activation
        # quantisation as its prologue, at decode shapes); under TP the row statistics are all-reduced
        cfg4_res = {}
        if not os.environ.get("Q100_BENCH_SKIP_CFG4"):
            for tag, wb, ab in (("w8a8", 8, 8), ("w4a4", 4, 4)):
                torch.cuda.empty_cache()
                cfg4_res[tag] = row(wbits=wb, abits=ab, group=0, ppl_samples=1, ppl_ref=True, batches="1,8", decode_steps=24)
                extra["llama3_8b_smoothquant_configs"] = dict(cfg4_res)
        extra["llama3_8b_smoothquant_configs"] = cfg4_res
        model70_res = None
        if not os.environ.get("Q100_BENCH_SKIP_70B"):
            torch.cuda.empty_cache()
            model70_res = row(model="70b", ppl_samples=1, ppl_ref=True, ppl_layers=4, batches="1,16", decode_steps=24)
            extra["llama3_70b"] = model70_res

    if world > 1:
        from q100 import peer as _peer
        _one = _peer.get(None)
        if _one is not None:
            _one.check()                                            # a timed-out peer all-reduce voids the decode numbers
    done_flag.set()
    wd.cancel()
    if rank == 0:
        line = make_line()
        print(json.dumps(line), flush=True)
    if world > 1:
        _teardown_distributed()


def _teardown_distributed():
    """Peer buffers (symmetric memory) must go before the process group; a watchdog makes sure a stuck teardown can
    never keep the job alive after the result line is out."""
    import gc
    import threading
    import torch.distributed as dist
    threading.Timer(45.0, lambda: os._exit(0)).start()
    try:
        from q100 import peer
        peer.shutdown()
    except Exception:                                               # noqa: BLE001
        pass
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
