"""The fp32 ranking of the nine part pairs in k_pair_eval's distance_three_circles (csrc/pair_kernels.cuh, PAIR_HMIN_RANKED) is
exact by construction; scripts/validate_hmin_ranking.c restates the selection with the kernel's constants and compares it bit
for bit with the reference's nine-fold loop (core/distance.py:55-105) on the host.  The constants are read out of the CUDA
source so that the two cannot drift apart unnoticed."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constants_match_the_kernel():
    cu = open(os.path.join(ROOT, 'crowddynamics_b200', 'csrc', 'pair_kernels.cuh')).read()
    c = open(os.path.join(ROOT, 'scripts', 'validate_hmin_ranking.c')).read()
    k = re.search(r'thr = ha_lo \+ 2\.0f \* \((1e-6f) \* \(d_hi \+ rit \+ ris \+ rjt \+ rjs\) \+ (1e-15f)\)', cu)
    h = re.search(r'thr=lo\+2\.0f\*\((1e-6f)\*\(dhi\+rif\[0\]\+rif\[1\]\+rjf\[0\]\+rjf\[1\]\)\+(1e-15f)\)', c)
    assert k and h and k.groups() == h.groups()
    assert 'fabsf(chk) < 1e30f' in cu and 'fabsf(chk)<1e30f' in c


def test_ranked_selection_equals_the_nine_fold_loop(tmp_path):
    exe = str(tmp_path / 'validate_hmin_ranking')
    subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-o', exe, os.path.join(ROOT, 'scripts', 'validate_hmin_ranking.c'), '-lm'])
    out = subprocess.run([exe, '300000'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout
    assert 'mismatches 0' in out.stdout
