"""TEST INFRASTRUCTURE ONLY.

`oracle/` is the CPU restatement of the reference rollout path (Flow-Factory SD3.5 adapter +
FlowMatchEulerDiscreteSDEScheduler + the vendored diffusers SD3Transformer2DModel forward).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it - as the checker or the timed CPU baseline, never as the product path.
The product (`flow_factory_b200`) never imports this package.

Parity pinning: the reference has NO golden vectors for this path; the oracle is
pinned against outputs of the reference itself, imported read-only in the build container by
`tests/golden/make_golden.py` (fixtures committed under tests/golden/*.pt).
"""
