"""The C++ host mirror (include/v2p_host.hpp): builds everywhere, its host-only part runs on CPU, the full program
(golden vectors through v2p::GIR::execute and v2p::HaplotypeBatch in three layouts) runs on the GPU."""
import os
import subprocess

import pytest

from oracle import taskgen
from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_test.cpp")


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    from vcf2prot_b200 import _lib

    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    out = str(tmp_path_factory.mktemp("cpp") / "host_test")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", out,
                           "-L", libdir, "-lv2p_engine", "-Wl,-rpath," + libdir])
    return out


def write_cases(path):
    n = 0
    with open(path, "w") as f:
        for case in load_golden("unit_tests.json"):
            if not case["tasks"]:
                continue
            refs = {case["transcript"]: case["ref"]}
            muts = taskgen.alt_transcript(case["transcript"], case["csqs"])
            g = taskgen.TranscriptInstruction.from_alt_transcript(case["transcript"], muts, refs).get_g_rep(refs)
            f.write("CASE %s\nREF %s\nALT %s\nRES %d\n" % (case["name"], g.ref, g.alt or "-", g.res_len))
            for t in case["tasks"]:  # the reference binary's own Vec<Task> dump
                f.write("TASK %d %d %d %d\n" % tuple(t))
            f.write("EXPECT %s\n" % (case["records"][0][1] or "-"))
            n += 1
    return n


def test_cpp_host_builds_and_host_only_part_passes(binary):
    p = subprocess.run([binary, "--no-gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0 and "OK" in p.stdout, p.stdout


@pytest.mark.gpu
def test_cpp_host_golden_vectors_on_gpu(binary, tmp_path):
    cases = str(tmp_path / "cases.txt")
    assert write_cases(cases) >= 25
    outdir = tmp_path / "fasta"
    outdir.mkdir()
    p = subprocess.run([binary, cases, str(outdir)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0 and "OK" in p.stdout, p.stdout
    assert sorted(f.name for f in outdir.iterdir()) == ["HG1.fasta", "HG2.fasta", "HG3.fasta"]
