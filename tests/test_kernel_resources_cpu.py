"""Static guards on the compiled headline kernels (no GPU needed: cuobjdump reads the in-tree library).

A run-time branch added to the edge-tile path of the fused CG kernels once pushed phase B into register
spills and cost 7 % of the headline number without any test noticing.  These checks pin what
the measured build has: no spill stack in the CG TMA kernels of non-periodic problems, the register budget
that lets two CTAs share an SM, and TMA + mbarrier instructions in their SASS."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyapes_b200", "lib", "libpyapes_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(CUOBJDUMP) and shutil.which("c++filt")),
                                reason="needs cuobjdump and c++filt")


def _resources():
    out = subprocess.run([CUOBJDUMP, "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    names, usage = [], {}
    lines = out.splitlines()
    for i, line in enumerate(lines):
        m = re.match(r"\s*Function (\S+):", line)
        if m and i + 1 < len(lines):
            r = re.search(r"REG:(\d+) STACK:(\d+)", lines[i + 1])
            if r:
                names.append(m.group(1))
                usage[m.group(1)] = (int(r.group(1)), int(r.group(2)))
    plain = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    return {p: (n, *usage[n]) for p, n in zip(plain, names)}


def test_headline_cg_kernels_do_not_spill():
    res = _resources()
    cg = {k: v for k, v in res.items() if re.search(r"k_cg_phase[AB]_tma<", k)}
    assert cg, "no fused CG TMA kernels in the library"
    # <T, tile kind, WRAP, UNI[, HALO]>: every non-periodic (WRAP = false) fp64 instantiation
    headline = {k: v for k, v in cg.items() if re.search(r"_tma<double, pa::K(Std|Flat), false, (true|false)", k)}
    assert len(headline) >= 6, sorted(cg)
    for name, (_, regs, stack) in headline.items():
        # the kernels are persistent: a few words of per-ITEM state (item id, pipeline counter, sums) may live
        # on the stack between items -- never inside the plane loops (checked below)
        assert stack <= 64, f"{name.split('(')[0]}: {stack} B of spill stack"
        assert regs <= 96, f"{name.split('(')[0]}: {regs} registers (two 288-thread CTAs per SM need <= 96)"


def test_headline_kernels_keep_local_memory_out_of_the_plane_loops():
    """The persistent kernels may park a few words of per-ITEM state (item id, pipeline counter, sums) on the
    stack between items: a handful of LDL / STL in the whole kernel.  Spilling inside a plane loop (3x unrolled,
    4 cells per thread, two paths) shows up as hundreds -- that is what this tripwire is for; the measured build
    has 25-45 per kernel, all in the item prologue / epilogue (cuobjdump -sass)."""
    res = _resources()
    checked = 0
    for name, (mangled, _, _) in res.items():
        if not re.search(r"k_cg_phase[AB]_tma<double, pa::KStd, false, (true|false)(, false)*>", name):
            continue
        sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", mangled, LIB], capture_output=True, text=True).stdout
        n = len(re.findall(r"\s(LDL|STL)[. ]", sass))
        assert n <= 64, f"{name.split('(')[0]}: {n} local-memory instructions"
        checked += 1
    assert checked >= 3


def test_headline_cg_kernels_use_tma_and_mbarriers():
    res = _resources()
    for pat in (r"k_cg_phaseA_tma<double, pa::KStd, false, false, false>", r"k_cg_phaseB_tma<double, pa::KStd, false, true, false, false>"):
        hits = [v[0] for k, v in res.items() if re.search(pat, k)]
        assert len(hits) == 1, pat
        sass = subprocess.run([CUOBJDUMP, "-sass", "-fun", hits[0], LIB], capture_output=True, text=True).stdout
        assert "UTMALDG" in sass, f"{pat}: no TMA tensor loads in the SASS"
        assert "SYNCS" in sass, f"{pat}: no mbarrier instructions in the SASS"


def test_resident_kernels_do_not_spill_and_fit_one_cta_of_512_threads():
    """The shared-memory-resident kernels (kernels_resident.cuh) run ONE 512-thread CTA per SM: <= 128 registers, and
    no spill stack -- a noinline call of the general cell path once pushed every item's vectors through local memory."""
    res = _resources()
    hits = {k: v for k, v in res.items() if re.search(r"k_(euler|cg)_resident<", k)}
    assert len(hits) >= 14, sorted(hits)
    for name, (_, regs, stack) in hits.items():
        assert stack == 0, f"{name.split('(')[0]}: {stack} B of spill stack"
        assert regs <= 128, f"{name.split('(')[0]}: {regs} registers"
