"""Static guard on the compiled half-sweep kernels (no GPU needed: cuobjdump reads liblqft.so).

ptxas moves non-coherent loads (ld.global.nc / LDG.E....CONSTANT) above griddepcontrol.wait (ACQBULK): with the wait
taken late in k_sweep_n16 that once put five field loads of a slab instantiation in front of it and the self-halo parity
test failed on the acceptance count (profiles/r2_late_wait.txt, item 4).  The loads right behind the wait are therefore
coherent loads in volatile statements (ld_after_wait, kernels_common.cuh).  This test keeps it that way: in every
instantiation of k_sweep_n16 no 128-bit global load and no store of the field may sit in front of the last wait."""
import re
import shutil
import subprocess

import pytest

from lattice_qft_b200 import _lib


@pytest.fixture(scope="module")
def sass_by_kernel():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=600)
    except (FileNotFoundError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    if out.returncode != 0:
        pytest.skip("cuobjdump failed: " + out.stderr[:200])
    kernels, name = {}, None
    for line in out.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            kernels[name].append((int(m.group(1), 16), m.group(2)))
    return kernels


def test_no_field_access_in_front_of_the_wait(sass_by_kernel):
    sweeps = {k: v for k, v in sass_by_kernel.items() if "k_sweep_n16" in k}
    assert len(sweeps) >= 100, f"expected every instantiation of k_sweep_n16, found {len(sweeps)}"
    for name, code in sweeps.items():
        waits = [a for a, op in code if "ACQBULK" in op]
        assert waits, f"{name}: no griddepcontrol.wait"
        last = max(waits)
        early = [(hex(a), op) for a, op in code if a < last and re.search(r"\b(LDG|LD)\.E\.128|\b(STG|ST)\.E\.128", op)]
        assert not early, f"{name}: field access in front of the wait at {hex(last)}: {early[:3]}"
        # the loads right behind the wait are the coherent ones
        behind = [op for a, op in code if a > last and re.search(r"\bLDG\.E\.128", op)][:6]
        assert len(behind) == 6 and not any("CONSTANT" in op for op in behind), f"{name}: {behind}"


def test_headline_kernel_keeps_its_code_shape(sass_by_kernel):
    """k_sweep_n16<16, plain, lone chain> (256^3): ptxas keeps the Philox round keys in uniform registers and the row loop
    free of divergence handling only while the control flow is visibly warp uniform; when it stops doing so
    the same source compiles to ~40 % more instructions and runs 18 % slower.  The numbers of the shipped build:
    1272 instructions, 416 of them on the uniform datapath, no BRA.DIV."""
    names = [k for k in sass_by_kernel if "k_sweep_n16ILi16ELb0ELb0ELb1ELb0ELb0E" in k]
    assert len(names) == 1, names
    code = sass_by_kernel[names[0]]
    ops = [op.split()[1] if op.startswith("@") else op.split()[0] for _, op in code]
    assert not any(o.startswith("BRA.DIV") for o in ops)
    uniform = sum(1 for (_, op), o in zip(code, ops) if o.startswith("U") or "UR" in op)
    assert len(ops) <= 1400, len(ops)
    assert uniform >= 350, uniform
