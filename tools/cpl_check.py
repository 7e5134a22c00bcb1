"""A/B of the one-launch coupling step against the three-launch chain (run once per BFSR_FUSE_CPL setting, then compare):
    BFSR_FUSE_CPL=0 python tools/cpl_check.py run gpurun_out/cpl_a.pt ; BFSR_FUSE_CPL=1 python tools/cpl_check.py run gpurun_out/cpl_b.pt
    python tools/cpl_check.py cmp gpurun_out/cpl_a.pt gpurun_out/cpl_b.pt"""
import sys

import torch

sys.path.insert(0, ".")


def run(path):
    from bfsr_b200 import _lib
    from tools import synth
    t = synth.SRFlowTopo(nb=1, blocks=(0, 0, 0, 0), K=1)
    sd = synth.synth_srflow_state_dict(t, seed=3)
    table, keep = _lib.tensor_table(sd)
    L = _lib.lib()
    import os
    prec = int(os.environ.get("BFSR_CPL_CHECK_PREC", "0"))
    out = {}
    for C, layer, shapes in ((12, 3, ((2, 24, 20), (1, 96, 80), (3, 61, 45), (2, 320, 320))), (24, 8, ((2, 24, 20), (3, 61, 45), (4, 160, 160)))):
        for (B, H, W) in shapes:
            g = torch.Generator().manual_seed(H * 1000 + W + C)
            z = torch.randn(B, C, H, W, generator=g).cuda()
            ft = (torch.randn(B, 320, H, W, generator=g) * 0.5).cuda()
            for rev in (0, 1):
                o = torch.empty_like(z)
                _lib.check(L.bfsr_op_flowstep(table, len(table), f"flowUpsamplerNet.layers.{layer}".encode(), C, 1, rev, z.data_ptr(), ft.data_ptr(),
                                              B, H, W, o.data_ptr(), prec, 1, None))
                torch.cuda.synchronize()
                out[f"C{C}_{B}x{H}x{W}_rev{rev}"] = o.cpu()
    torch.save(out, path)
    print("saved", path, {k: float(v.abs().mean()) for k, v in out.items()})


def cmp(pa, pb, tol=2e-5):
    a, b = torch.load(pa), torch.load(pb)
    bad = 0
    for k in a:
        d = (a[k].double() - b[k].double())
        rel = float(d.norm() / a[k].double().norm())
        mx = float(d.abs().max())
        print(f"{k:20s} rel-L2 {rel:.3e} max-abs {mx:.3e} finite {bool(torch.isfinite(b[k]).all())}")
        if not (rel < tol) or not bool(torch.isfinite(b[k]).all()):
            bad += 1
            idx = (d.abs() > 1e-3).nonzero()
            print("   first bad indices:", idx[:8].tolist(), " count", idx.shape[0])
    print("CMP", "FAIL" if bad else "OK")
    return bad


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        sys.exit(1 if cmp(sys.argv[2], sys.argv[3]) else 0)
