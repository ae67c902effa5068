"""Guards on the generated code of the hot path (no GPU needed: cuobjdump / nvdisasm on the built library).

The fast interval is one ~1,000-instruction basic block; ptxas' schedule of it, and whether it spills inside it,
changed more than once in round 1 with edits to code that never runs in that block.  The in-order
issue model of profiles/tools/hot_block_model.py tracked the measured effect, so it is asserted here with a margin."""
import importlib.util
import os
import re
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hot_blocks(built_lib, pkg):
    if not (shutil.which("cuobjdump") and shutil.which("nvdisasm")):
        pytest.skip("CUDA binary utilities not installed")
    spec = importlib.util.spec_from_file_location("hot_block_model", os.path.join(ROOT, "profiles", "tools", "hot_block_model.py"))
    hbm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(hbm)
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", pkg._lib.LIB_PATH], cwd=d, check=True, capture_output=True)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", os.path.join(d, cubin)], capture_output=True, text=True, check=True).stdout.split("\n")
    out = {}
    for l in sass:
        m = re.match(r"^\.text\.(_Z\d+rsf_(mcmc|forward|init)\w*):$", l)
        if m:
            hot = max(hbm.blocks_of(sass, m.group(1)), key=len)
            out[m.group(1)] = (len(hot), hbm.model_cycles(hot), sum(1 for i in hot if re.search(r"\b(LDL|STL)", i)))
    return out


def _one(hot_blocks, prefix):
    ks = [k for k in hot_blocks if k.startswith(prefix)]
    assert len(ks) == 1, (prefix, ks)
    return hot_blocks[ks[0]]


@pytest.mark.parametrize("prefix, max_cycles", [
    ("_Z20rsf_mcmc_spec_kernelILi1ELb0ELb0EE", 2020),      # bench config (cfg 2): 1,924 in the round-1 build
    ("_Z15rsf_mcmc_kernelILi1ELb0ELb0ELb0ELb0EE", 2180),           # saturating sizes / cfg 5: 2,072
    ("_Z15rsf_mcmc_kernelILi3ELb0ELb0ELb0ELb0EE", 2160),          # cfg 3 without round packing: 2,052
    ("_Z15rsf_mcmc_kernelILi3ELb0ELb0ELb1ELb0EE", 2160),           # cfg 3: 2,052
    ("_Z18rsf_forward_kernelILi1ELb0ELb0EE", 2000),            # forward batches: 1,903
])
def test_fast_interval_block_schedule_and_no_spills(hot_blocks, prefix, max_cycles):
    n, cycles, local = _one(hot_blocks, prefix)
    assert 900 <= n <= 1100            # it is the fast interval (probe + twelve stages + error forms), in one block
    assert local == 0                  # no spill inside it
    assert cycles <= max_cycles


def test_both_variants_of_every_solver_kernel_are_built(hot_blocks):
    for stem in ("rsf_forward_kernelILi1E", "rsf_init_kernelILi1E", "rsf_init_kernelILi3E",
                 "rsf_mcmc_kernelILi1ELb0E", "rsf_mcmc_kernelILi3ELb0E", "rsf_mcmc_spec_kernelILi1ELb0E"):
        # (the template arguments are ..., VS[, PACK], K1P, except for the speculative kernel, which ends in VS)
        names = []
        for n in (k.split("EEv")[0] for k in hot_blocks if stem in k):
            if "rsf_mcmc_spec_kernelI" not in n:
                if n.endswith("ELb1"):
                    continue                                   # the k1 instantiation (d = 1, default variant only)
                n = n[:-4]
            if "rsf_mcmc_kernelI" in n:
                if n.endswith("ELb1"):
                    continue                                   # the packing instantiation (d = 3, default variant only)
                n = n[:-4]
            names.append(n)
        names.sort()
        assert len(names) == 2 and names[0].endswith("Lb0") and names[1].endswith("Lb1"), (stem, names)


def test_k1_instantiations_are_built_beside_the_tuned_kernels(hot_blocks):
    """RSFM_PARAM_K1 lives in its own instantiations (forward, init, the d = 1 one-thread-per-chain kernel with Philox
    and with host-supplied draws), so the per-chain k1 register never enters the tuned kernels."""
    for name in ("_Z18rsf_forward_kernelILi1ELb0ELb1EE", "_Z15rsf_init_kernelILi1ELb0ELb1EE",
                 "_Z15rsf_mcmc_kernelILi1ELb0ELb0ELb0ELb1EE", "_Z15rsf_mcmc_kernelILi1ELb1ELb0ELb0ELb1EE"):
        n, cycles, local = _one(hot_blocks, name)
        assert 900 <= n <= 1100 and local == 0
