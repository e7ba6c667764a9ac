"""CPU suite: the device task generator's emission rules (csrc/v2p_taskgen_rules.cuh, compiled here for the host) against
the reference-pinned oracle (oracle/taskgen.py) on the reference's own unit-test inputs, the golden combos and seeded
random mutation sets over all 22 supported consequence classes -- tasks, alteration bytes, result size and the
skip / abort outcomes (transcript_instructions.rs:41-63, :214-321, :335-780; instruction.rs:1075-1098)."""
import os
import random
import subprocess

import pytest

from oracle import taskgen as T
from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OK, EMPTY, ABSENT, SKIPPED, PANIC = range(5)
AA = "ACDEFGHIKLMNPQRSTVWY"


def static_instruction(m):
    """What the HOST hands to the device per mutation: the instruction assuming validate_s_state passes, plus the two
    context flags the device needs to redo that validation per haplotype."""
    ins = T.instruction_from_mutation(m, [m])
    star = m.mut_type.startswith("*")
    inval = m.mut_type in ("stop_gained", "frameshift", "*stop_gained") or (
        m.mut_type in ("inframe_insertion", "inframe_deletion") and m.mut_aa[0] in ("Not", "End"))
    return ins, (1 if star else 0) | (2 if inval else 0)


def oracle_outcome(name, muts, refs):
    """-> (status, size, tasks, alt) exactly as haplotype_g_rep would treat this transcript."""
    try:
        ti = T.TranscriptInstruction.from_alt_transcript(name, muts, refs)
    except T.TaskGenError:
        return ABSENT, 0, [], ""
    try:
        size = ti.expected_results_size()
    except T.RefPanic:
        return PANIC, 0, [], ""
    try:
        g = ti.get_g_rep(refs)
    except T.TaskGenError:
        return SKIPPED, size, [], ""
    except T.RefPanic:
        return PANIC, size, [], ""
    if not g.tasks and g.res_len == 0 and any(i.code in "0U" for i in ti.instructions):
        return EMPTY, size, [], ""
    return OK, size, g.tasks, g.alt


def case_text(name, muts, ref_len):
    muts = sorted(muts, key=lambda m: m.mut_pos)
    lines = ["CASE %s %d %d" % (name, ref_len, len(muts))]
    for m in muts:
        ins, flags = static_instruction(m)
        lines.append("INS %s %d %d %d %d %s" % (ins.code, flags, ins.pos_ref, ins.pos_res, ins.len, ins.data or "-"))
    return "\n".join(lines) + "\n"


@pytest.fixture(scope="module")
def rules_binary(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("rules") / "taskgen_rules_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "vcf2prot_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "taskgen_rules_test.cpp"), "-o", out])
    return out


def run_cases(binary, cases):
    """cases: [(name, muts, refs)] -> compares every field with the oracle; returns the status histogram."""
    text = "".join(case_text(n, m, len(r[n])) for n, m, r in cases)
    p = subprocess.run([binary], input=text, stdout=subprocess.PIPE, text=True, check=True)
    blocks, cur = {}, None
    for line in p.stdout.splitlines():
        f = line.split()
        assert f[0] != "MISMATCH", line
        if f[0] == "RESULT":
            cur = blocks[f[1]] = {"status": int(f[2]), "size": int(f[3]), "n_tasks": int(f[4]), "n_alt": int(f[5]), "tasks": [], "alt": ""}
        elif f[0] == "TASK":
            cur["tasks"].append(tuple(int(x) for x in f[1:]))
        elif f[0] == "ALT":
            cur["alt"] = "" if f[1] == "-" else f[1]
    hist = [0] * 5
    for name, muts, refs in cases:
        st, size, tasks, alt = oracle_outcome(name, muts, refs)
        got = blocks[name]
        hist[st] += 1
        assert got["status"] == st, (name, got, st)
        if st in (OK, EMPTY, SKIPPED):
            assert got["size"] == size, (name, got, size)
        if st == OK:
            assert got["tasks"] == [tuple(t) for t in tasks], (name, got["tasks"], tasks)
            assert got["alt"] == alt and got["n_tasks"] == len(tasks) and got["n_alt"] == len(alt), name
    return hist


def test_reference_unit_tests_and_combos(rules_binary):
    cases = []
    for c in load_golden("unit_tests.json"):
        refs = {c["transcript"]: c["ref"]}
        cases.append((c["transcript"] + "." + c["name"].replace(" ", "_"), T.alt_transcript(c["transcript"], c["csqs"]), refs))
    for k, c in enumerate(load_golden("combos.json")):
        for name, muts in T.group_muts_per_transcript(c["csqs"]):
            if name in c["refs"]:
                cases.append(("%s.combo%d" % (name, k), muts, c["refs"]))
    # names must match the refs dict key: rebuild with unique keys
    fixed = []
    for i, (nm, muts, refs) in enumerate(cases):
        key = "C%d" % i
        tname = nm.split(".")[0]
        fixed.append((key, muts, {key: refs[tname]}))
    hist = run_cases(rules_binary, fixed)
    assert hist[OK] >= 25 and hist[EMPTY] >= 1 and hist[SKIPPED] >= 1


def random_csq(rng, tname, L):
    """One csq string of a random supported class with plausible (and sometimes implausible) fields."""
    typ = rng.choice(T.SUP_TYPE)
    pos = rng.randint(2, L)
    seq = lambda n: "".join(rng.choice(AA) for _ in range(n))
    star = lambda s: s + "*" if rng.random() < 0.3 else s
    base = typ.lstrip("*").split("&")[0]
    if base == "start_lost":
        aa = "1M>1%s" % seq(1)
    elif base == "missense":
        k = rng.choice([1, 1, 1, 2, 4])
        aa = "%d%s>%d%s" % (pos, seq(k), pos, star(seq(rng.choice([k, k, k + 1]))))
    elif base == "stop_gained":
        aa = "%d%s>%d*" % (pos, seq(rng.choice([1, 1, 3])), pos)
    elif base == "stop_lost":
        p = rng.choice([L, L + 1, pos])
        aa = "%d*>%d%s" % (p, p, star(seq(rng.randint(1, 12))))
    elif base == "frameshift":
        aa = "%d%s>%d%s" % (pos, star(seq(rng.randint(1, 6))), pos, rng.choice([star(seq(rng.randint(1, 20))), "*"]))
    elif base == "inframe_insertion":
        aa = "%d%s>%d%s" % (pos, star(seq(rng.choice([1, 1, 1, 2]))), pos, rng.choice([star(seq(rng.randint(2, 8))), "*"]))
    else:  # inframe_deletion
        k = rng.randint(2, 7)
        aa = "%d%s>%d%s" % (pos, star(seq(k)), pos, rng.choice([seq(1), seq(1), star(seq(rng.randint(1, 3))), "*"]))
    return "%s|GENE|%s|protein_coding|-|%s|1C>T" % (typ, tname, aa)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_mutation_sets_over_all_classes(rules_binary, seed):
    rng = random.Random(seed)
    cases = []
    for i in range(1500):
        L = rng.randint(12, 90)
        key = "R%d" % i
        refs = {key: "M" + "".join(rng.choice(AA) for _ in range(L - 1))}
        csqs, used = [], set()
        for _ in range(rng.choice([1, 1, 2, 2, 3, 4])):
            c = random_csq(rng, key, L)
            try:
                m = T.Mutation.from_csq(c)
                static_instruction(m)  # the host-side interpretation itself may abort the reference: not a device matter
            except (T.TaskGenError, T.RefPanic):
                continue
            if m.mut_pos in used:  # duplicates are dropped / abort before instruction generation (vcf_ds.rs:442-479)
                continue
            used.add(m.mut_pos)
            csqs.append(c)
        if csqs:
            cases.append((key, T.alt_transcript(key, csqs), refs))
    hist = run_cases(rules_binary, cases)
    assert hist[OK] > 300 and hist[SKIPPED] > 20 and hist[ABSENT] > 5, hist


def test_reference_task_builder_unit_tests(rules_binary):
    """transcript_instructions.rs:806-882 (test_get_task_from_frameshift / _stop_gained / _stop_lost / _inframe_insersion):
    the exact Task tuples and alt-tape contents those tests assert, produced here from whole transcripts (so the frameshift
    task starts where ITS base task ends, 39, not at the hand-set 30 of the unit test)."""
    def run(csq_aa, typ, ref_len):
        name = "ENST00000000001"
        m = T.Mutation.from_csq("%s|GENE|%s|protein_coding|-|%s|1C>T" % (typ, name, csq_aa))
        p = subprocess.run([rules_binary], input=case_text(name, [m], ref_len), stdout=subprocess.PIPE, text=True, check=True)
        tasks = [tuple(int(x) for x in l.split()[1:]) for l in p.stdout.splitlines() if l.startswith("TASK")]
        alt = [l.split()[1] for l in p.stdout.splitlines() if l.startswith("ALT")]
        return tasks, (alt[0] if alt and alt[0] != "-" else "")

    tasks, alt = run("40VGLHFWTM*>40VDSTFGQC", "frameshift", 48)
    assert tasks == [(0, 0, 39, 0), (1, 0, 8, 39)] and alt == "VDSTFGQC"  # :806-822  Task::new(1, 0, 8, <end of the base task>)
    tasks, alt = run("40VGLHFWTM*>40*", "stop_gained", 48)
    assert tasks == [(0, 0, 39, 0)] and alt == ""                        # :824-840  phi (2,0,0,0): nothing is pushed
    tasks, alt = run("489*>489S", "stop_lost", 488)
    assert tasks == [(0, 0, 488, 0), (1, 0, 1, 488)] and alt == "S"      # :842-861
    tasks, alt = run("125Y>125YRR", "inframe_insertion", 300)
    assert tasks[:2] == [(0, 0, 124, 0), (1, 0, 3, 124)] and alt == "YRR"  # :863-882
    assert tasks[2] == (0, 125, 175, 127)
