"""Fused (nqb_tp_fused_fwd) vs unfused (k_hidden_fwd + k_gemm3x + tp_fwd*) forward of every layer of a bench model,
timed with CUDA events on the launching stream.  python tools/bench_fused.py [--workload li3po4_10k_l2_f64]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import R_MAX, WORKLOADS, build_system  # noqa: E402
from nequip_b200 import data as D, ops  # noqa: E402
from nequip_b200.nn import dense  # noqa: E402
from nequip_b200.nn.model import NequIPEnergyModel, ScalarLinearLayer  # noqa: E402


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="li3po4_10k_l2_f64")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--layers", default="")
    ap.add_argument("--prof", action="store_true", help="per-role stall counters of CTA 0 (generated with fused_prof)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    sysd, meta, mk = build_system(args.workload, seed=0)
    model = NequIPEnergyModel(r_max=R_MAX, type_names=meta["type_names"], parity=True,
                              avg_num_neighbors=meta["avg_num_neighbors"], **mk).to(dev)
    for p in model.parameters():
        p.requires_grad_(False)
    ei = sysd["edge_index"].to(dev)
    N, E = sysd["pos"].shape[0], ei.shape[1]
    csr = ops.build_csr(ei[0].contiguous(), N)
    src = ei[1].contiguous()
    g = torch.Generator(device=dev).manual_seed(0)
    want = [int(v) for v in args.layers.split(",")] if args.layers else range(len(model.layers))
    for li in want:
        conv = model.layers[li].conv
        plan = conv.tp_scatter._plan
        sig = plan.sig
        lins = [m for m in conv.edge_mlp.mlp if isinstance(m, ScalarLinearLayer)]
        x = torch.randn(N, sig.d_in, device=dev, generator=g)
        y = torch.randn(E, sig.s_dim, device=dev, generator=g)
        emb = torch.rand(E, 8, device=dev, generator=g)
        mlp = dense.RadialMLPGemm(lins[0], lins[1], dev)
        with torch.no_grad():
            def unfused():
                w = mlp(emb)
                return ops.tp_scatter(plan, x, y, w, ei[0], src, csr=csr)

            ms_u = timeit(unfused, args.reps)
            row = {"layer": li, "W": sig.weight_numel, "paths": len(sig.paths), "N": N, "E": E, "unfused_ms": round(ms_u, 4)}
            if dense.FusedRadialTP.supported(lins[0], lins[1], plan, torch.float32):
                fz = dense.FusedRadialTP(lins[0], lins[1], plan, dev)
                h = torch.empty(E, 128, device=dev)

                def fused(want_w):
                    ops.mlp_hidden_fwd(emb, fz.w1s, h, None)
                    return ops.tp_fused_fwd(fz.fw, x, y, h, src, csr, want_w=want_w)

                ms_f = timeit(lambda: fused(False), args.reps)
                ms_fw = timeit(lambda: fused(True), args.reps)
                hid = timeit(lambda: ops.mlp_hidden_fwd(emb, fz.w1s, h, None), args.reps)
                ref = unfused()
                out, _ = fused(False)
                err = float((out - ref).abs().max() / ref.abs().max())
                flops = 2.0 * E * 128 * sig.weight_numel
                row.update({"fused_ms": round(ms_f, 4), "fused_with_w_out_ms": round(ms_fw, 4), "hidden_ms": round(hid, 4),
                            "slices": fz.fw.nslice, "ctas": fz.fw.nctas, "rel_diff_vs_unfused": err,
                            "tensor_TFLOPs_3xtf32": round(3 * flops / ((ms_f - hid) * 1e-3) / 1e12, 1)})
                if args.prof:
                    import ctypes
                    import dataclasses

                    popts = dataclasses.replace(plan.opts, fused_prof=True)
                    pplan = ops.get_plan(sig.irreps_in1, sig.irreps_in2, sig.irreps_out, sig.instructions, popts)
                    pfw = ops.FusedTPWeights(pplan, lins[1].weight.detach(), float(lins[1].alpha), dev)
                    ops.tp_fused_fwd(pfw, x, y, h, src, csr, want_w=False)
                    torch.cuda.synchronize()
                    lib = ctypes.CDLL(pplan.spec_path)
                    buf = (ctypes.c_ulonglong * (160 * 32))()
                    lib.nqb_spec_fused_prof(buf)
                    v = list(buf)
                    per_slice = []
                    for si in range(fz.fw.nslice):
                        c = pfw.cta0[si]  # first CTA of the slice
                        r = v[c * 32:(c + 1) * 32]
                        tiles = max(1, r[7])
                        per_slice.append({
                            "slice": si, "ctas": pfw.cta0[si + 1] - c, "tiles": r[7],
                            "consumer/tile": {"acc_full": r[0] // tiles, "x_full": r[1] // tiles, "tmem_ld": r[2] // tiles,
                                              "store": r[3] // tiles, "total": r[6] // tiles},
                            "h_producer/tile": {"a_done": r[8] // tiles, "cp_wait": r[10] // tiles,
                                                "lo_pass": r[11] // tiles, "issue_h": r[13] // tiles, "total": r[14] // tiles},
                            "mma/tile": {"acc_empty": r[16] // tiles, "a_full": r[17] // tiles, "total": r[22] // tiles},
                            "stager/tile": {"x_empty": r[25] // tiles, "total": r[30] // tiles}})
                    row["prof_first_cta_of_slice"] = per_slice
                    row["cta_total_Mcycles"] = [round(v[c * 32 + 6] / 1e6, 2) for c in range(pfw.nctas)]
            print(json.dumps(row), flush=True)
        del x, y, emb


if __name__ == "__main__":
    main()
