"""Row N2 (as far as this environment allows: no rustc): the Rust sys crate is GENERATED from include/euc_b200.h, so it cannot
drift from the header.  The committed file must be what the generator produces now, and it must declare every export."""
import os
import re
import subprocess
import sys

from euc_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_RS = os.path.join(ROOT, "bindings", "rust", "euc-b200-sys", "src", "lib.rs")


def test_sys_crate_is_current():
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"]).returncode
    assert rc == 0, "bindings/rust/euc-b200-sys/src/lib.rs is stale: run python tools/gen_rust_sys.py"


def test_sys_crate_declares_every_export_and_struct():
    src = open(LIB_RS).read()
    declared = set(re.findall(r"pub fn (euc_\w+)\(", src))
    assert declared == set(abi.SYMBOLS), sorted(set(abi.SYMBOLS) ^ declared)
    for s in ("euc_sampler_desc", "euc_pipeline_desc", "euc_batch_draw", "euc_render_stats", "euc_uniforms_teapot_phong", "euc_vertex_voxel"):
        assert f"pub struct {s} " in src
    # the safe wrappers only call functions that exist
    wrap = open(os.path.join(ROOT, "bindings", "rust", "euc-b200", "src", "lib.rs")).read()
    used = set(re.findall(r"sys::(euc_\w+)\(", wrap))
    assert used and used <= declared, sorted(used - declared)
