"""Helpers for the -m gpu parity tests: call the C ABI through ctypes on torch-owned device memory."""
import ctypes as C
import json
import os

import torch

from flow_factory_b200 import _lib

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


def dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(obj, f, indent=1, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o))


def record(file: str, obj):
    """Append one JSON line to gpurun_out/<file>: the MEASURED error of every parity case, so tolerances can be set from numbers."""
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, file), "a") as f:
        f.write(json.dumps(obj, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o)) + "\n")


def err_report(got: torch.Tensor, ref: torch.Tensor, tag: str):
    g, r = got.float(), ref.float()
    d = (g - r).abs()
    rep = dict(tag=tag, shape=list(g.shape), max_abs=float(d.max()), mean_abs=float(d.mean()), ref_absmax=float(r.abs().max()),
               got_absmax=float(g.abs().max()), n_nan=int(torch.isnan(g).sum()),
               frac_bad=float((d > 0.05 * r.abs().max()).float().mean()))
    record("parity_measured.jsonl", {k: rep[k] for k in ("tag", "max_abs", "mean_abs", "ref_absmax", "n_nan")})
    if g.dim() == 2:
        bad = d > 0.05 * r.abs().max()
        rep["bad_rows_first"] = bad.any(1).nonzero().flatten()[:16].tolist()
        rep["bad_cols_first"] = bad.any(0).nonzero().flatten()[:16].tolist()
        rep["got_00"] = g[:2, :8].tolist()
        rep["ref_00"] = r[:2, :8].tolist()
    return rep


from flow_factory_b200.ops import linear  # noqa: E402,F401


def device_error():
    buf = (C.c_uint * 4)()
    _lib.lib().ffb200_device_error(C.byref(buf))
    return [hex(x) for x in buf]
