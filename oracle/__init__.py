"""TEST INFRASTRUCTURE ONLY -- parity oracle for stylerenderer_b200.

`oracle/` holds a CPU restatement of the reference hot path (plain C in sr_oracle.c /
raster_body.inc, torch-CPU module restatement in torch_ref.py) plus the recipe that compiles the
unmodified reference extensions into `oracle/_ref/` (build_ref.py).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this package.  `stylerenderer_b200` never does.
"""
