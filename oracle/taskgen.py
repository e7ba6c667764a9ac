"""ORACLE (test infrastructure only -- never imported by the product path).

Pure-Python restatement of the *producer* and *consumer* that sit either side of
vcf2prot's sequence-generation engine, so that parity tests can build reference-valid
Task arrays and slice the result tape exactly as the reference does.

Restated reference code (paths relative to /root/reference/src):
  functions/text_parser.rs:21-59        split_csq_string
  functions/text_parser.rs:83-139       parse_amino_acid_field / parse_amino_acid_seq_position
  data_structures/mutation_ds.rs:16-142 MutationType / MutatedString / MutationInfo / Mutation
  data_structures/vcf_ds.rs:358-441     AltTranscript::new / sort_alterations
  data_structures/InternalRep/instruction.rs:64-1098          Mutation -> Instruction (22 csq classes)
  data_structures/InternalRep/transcript_instructions.rs:41-169   from_alt_transcript
  data_structures/InternalRep/transcript_instructions.rs:214-321  compute_expected_results_array_size
  data_structures/InternalRep/transcript_instructions.rs:335-427  get_g_rep
  data_structures/InternalRep/transcript_instructions.rs:452-780  to_task and the per-class task builders
  data_structures/InternalRep/haplotype_instruction.rs:37-158     from_vec_t_ins / get_g_rep / update_task
  data_structures/InternalRep/personalized_genome.rs:61-117       from_proband_instruction / write_altered_only
  data_structures/InternalRep/sequence_tape.rs:33-89              SequenceTape::new / get_seq
  functions/vcf_tools.rs:82-133         group_muts_per_transcript / get_unique_transcript

Parity status: pinned.  tests/test_oracle_golden.py checks this restatement against
tests/golden/*.json, which oracle/make_golden.py harvested from the reference's own
prebuilt binary (oracle/_ref/vcf2prot, v0.1.2) replaying the reference unit-test inputs
(transcript_instructions.rs:884-1594) and seeded synthetic cohorts.

A Task here is the tuple (exe_code, start_pos, length, start_pos_res) -- task.rs:2-9.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

Task = Tuple[int, int, int, int]
PHI: Task = (2, 0, 0, 0)

SUP_TYPE = (  # Constants.rs:3-8
    "missense", "*missense", "frameshift", "*frameshift",
    "inframe_insertion", "*inframe_insertion", "inframe_deletion", "*inframe_deletion",
    "stop_gained", "stop_lost", "*missense&inframe_altering", "*frameshift&stop_retained",
    "*stop_gained&inframe_altering", "frameshift&stop_retained", "inframe_deletion&stop_retained",
    "inframe_insertion&stop_retained", "stop_gained&inframe_altering", "start_lost", "*stop_gained",
    "stop_lost&frameshift", "missense&inframe_altering", "start_lost&splice_region",
)


class TaskGenError(Exception):
    """Raised where the reference returns Err(..) (transcript is skipped by the caller)."""


class RefPanic(Exception):
    """Raised where the reference panics (process abort)."""


# --------------------------------------------------------------------------- parsing
def split_csq_string(s: str) -> List[str]:
    """text_parser.rs:21-59."""
    n = s.count("|")
    res = s.split("|")
    if n == 6:
        if res[3] in ("protein_coding", "NMD"):
            return [res[0], res[2], res[5]]
        raise TaskGenError("not a protein coding transcript")
    if res[0] == "start_lost":
        return [res[0], res[2], "1M>1*"]
    raise TaskGenError("incorrect number of fields")


def parse_amino_acid_seq_position(tok: str) -> Tuple[int, str]:
    """text_parser.rs:112-139 -- digits anywhere form the position, the rest is the sequence."""
    if "-" in tok:
        raise TaskGenError("'-' in amino acid field")
    digits = "".join(c for c in tok if c in "0123456789")
    if digits == "":
        raise TaskGenError("no position")
    pos = int(digits)
    if pos > 0xFFFF:  # u16 parse
        raise TaskGenError("position does not fit u16")
    seq = "".join(c for c in tok if c not in "0123456789")
    if seq == "":
        seq = "*"
    return pos, seq


def _mutated_string(s: str) -> Tuple[str, str]:
    """mutation_ds.rs:79-101 -> (kind, text), kind in {'Seq','End','Not'}."""
    if s == "":
        raise RefPanic("empty MutatedString")
    if s == "*":
        return ("Not", "*")
    if "*" in s:
        return ("End", s)
    return ("Seq", s)


@dataclass
class Mutation:
    transcript_name: str
    mut_type: str
    ref_pos: int  # 0-based (mutation_ds.rs:133-142)
    mut_pos: int
    ref_aa: Tuple[str, str]
    mut_aa: Tuple[str, str]

    @staticmethod
    def from_csq(csq: str) -> "Mutation":
        info = split_csq_string(csq)
        if info[0] not in SUP_TYPE:
            raise TaskGenError("unsupported mutation type %s" % info[0])
        parts = info[2].split(">")
        if len(parts) != 2:
            raise TaskGenError("bad aa field")
        rp, rs = parse_amino_acid_seq_position(parts[0])
        mp, ms = parse_amino_acid_seq_position(parts[1])
        if rp == 0 or mp == 0:
            raise RefPanic("u16 underflow on position 0")
        return Mutation(info[1], info[0], rp - 1, mp - 1, _mutated_string(rs), _mutated_string(ms))


def alt_transcript(name: str, csqs: Sequence[str]) -> List[Mutation]:
    """vcf_ds.rs:366-373 AltTranscript::new (unparsable csq strings are dropped)."""
    out = []
    for c in csqs:
        try:
            out.append(Mutation.from_csq(c))
        except TaskGenError:
            pass
    return out


# --------------------------------------------------------------------------- instructions
@dataclass
class Instruction:
    code: str
    s_state: bool
    pos_ref: int
    pos_res: int
    len: int
    data: str

    def key(self):
        return (self.code, self.s_state, self.pos_ref, self.pos_res, self.len, self.data)

    def __eq__(self, other):  # derived PartialEq (instruction.rs:6)
        return self.key() == other.key()


def _phi_ins() -> Instruction:
    return Instruction("E", False, 0, 0, 0, "")


def _strip_end(kind_text: Tuple[str, str]) -> str:
    """'data.remove(data.len()-1)' on an EndSequence: drops the LAST char whatever it is."""
    return kind_text[1][:-1]


def _validate_s_state(m: Mutation, muts: List[Mutation]) -> bool:
    """instruction.rs:1075-1098.  Mutation equality is mut_aa_position only (mutation_ds.rs:188-194)."""
    index = next(i for i, e in enumerate(muts) if e.mut_pos == m.mut_pos)
    for e in muts[:index]:
        if e.mut_type in ("stop_gained", "frameshift", "*stop_gained"):
            return False
        if e.mut_type in ("inframe_insertion", "inframe_deletion") and e.mut_aa[0] in ("Not", "End"):
            return False
    return True


def _i_missense(m, muts):  # instruction.rs:183-203
    k = m.mut_aa[0]
    if k == "Seq":
        data = m.mut_aa[1]
    elif k == "End":
        data = _strip_end(m.mut_aa)
    else:
        raise RefPanic("missense with '*' mutated aa")
    return Instruction("M", False, m.ref_pos, m.mut_pos, 1, data)


def _i_s_missense(m, muts):  # :222-238
    if _validate_s_state(m, muts):
        i = _i_missense(m, muts)
        i.code, i.s_state = "N", True
        return i
    return _phi_ins()


def _code2or3(m) -> Instruction:
    """Shared '2'/'3' arm (instruction.rs:270-306, :417-453, :1008-1044): note pos_ref/pos_res swapped."""
    k = m.mut_aa[0]
    data = m.mut_aa[1] if k == "Seq" else _strip_end(m.mut_aa)
    rk = m.ref_aa[0]
    ref_seq = m.ref_aa[1] if rk == "Seq" else _strip_end(m.ref_aa)
    pos_res, pos_ref = m.ref_pos, m.mut_pos
    if len(data) != len(ref_seq):
        return Instruction("3", False, pos_ref, pos_res, len(ref_seq), data)
    return Instruction("2", False, pos_ref, pos_res, len(data), data)


def _i_stop_gained(m, muts):  # :609-618
    return Instruction("G", False, m.ref_pos, m.mut_pos, 0, "")


def _i_stop_lost(m, muts):  # :672-691
    k = m.mut_aa[0]
    if k == "Seq":
        data = m.mut_aa[1]
    elif k == "End":
        data = _strip_end(m.mut_aa)
    else:
        raise RefPanic("stop_lost with '*' mutated aa")
    return Instruction("L", False, m.ref_pos, m.mut_pos, len(data), data)


def _i_frameshift(m, muts):  # :528-548
    k = m.mut_aa[0]
    if k == "Seq":
        data = m.mut_aa[1]
    elif k == "End":
        data = _strip_end(m.mut_aa)
    else:
        return _phi_ins()
    return Instruction("F", False, m.ref_pos, m.mut_pos, len(data), data)


def _i_inframe_insertion(m, muts):  # :257-328
    rk = m.ref_aa[0]
    if rk == "Seq":
        if len(m.ref_aa[1]) != 1:
            if m.mut_aa[0] == "Not":
                return _i_stop_gained(m, muts)
            # ref_aa is a Sequence here so its NotSeq arm (:291) is unreachable
            return _code2or3(m)
    elif rk == "End":
        return _i_frameshift(m, muts)
    else:
        raise RefPanic("inframe insertion with '*' reference")
    k = m.mut_aa[0]
    if k == "Seq":
        data = m.mut_aa[1]
    elif k == "End":
        return _i_frameshift(m, muts)
    else:
        return _i_stop_gained(m, muts)
    return Instruction("I", False, m.ref_pos, m.mut_pos, len(data), data)


def _i_s_inframe_insertion(m, muts):  # :347-370
    if _validate_s_state(m, muts):
        i = _i_inframe_insertion(m, muts)
        if i.code == "I":
            i.code, i.s_state = "J", True
        return i
    return _phi_ins()


def _i_inframe_deletion(m, muts):  # :389-474
    rk = m.ref_aa[0]
    if rk == "Seq":
        ln = len(m.ref_aa[1])
    elif rk == "End":
        ln = len(_strip_end(m.ref_aa))
    else:
        return _i_stop_gained(m, muts)
    k = m.mut_aa[0]
    if k == "Seq":
        if len(m.mut_aa[1]) == 1:
            data = m.mut_aa[1]
        else:
            if rk == "Not":
                raise RefPanic("unreachable")
            return _code2or3(m)
    elif k == "End":
        data = _strip_end(m.mut_aa)
        if len(data) != 1:
            return _i_frameshift(m, muts)
    else:
        return _i_stop_gained(m, muts)
    return Instruction("D", False, m.ref_pos, m.mut_pos, ln - len(data), data)


def _i_s_inframe_deletion(m, muts):  # :493-509  (code forced to 'C' whatever the inner arm returned)
    if _validate_s_state(m, muts):
        i = _i_inframe_deletion(m, muts)
        i.code, i.s_state = "C", True
        return i
    return _phi_ins()


def _i_s_frameshift(m, muts):  # :567-590
    if _validate_s_state(m, muts):
        if m.mut_aa[0] == "Not":
            return _i_stop_gained(m, muts)
        i = _i_frameshift(m, muts)
        i.code, i.s_state = "R", True
        return i
    return _phi_ins()


def _i_s_stop_gained(m, muts):  # :637-653
    if _validate_s_state(m, muts):
        i = _i_stop_gained(m, muts)
        i.code, i.s_state = "X", True
        return i
    return _phi_ins()


def _i_start_lost(m, muts):  # :710-719
    return Instruction("0", False, 0, 0, 0, "")


def _recode(i: Instruction, code: str) -> Instruction:
    if i.code != "E":
        i.code = code
    return i


def _i_s_frameshift_and_stop_retained(m, muts):  # :766-795
    if m.mut_aa[0] == "Not":
        if _validate_s_state(m, muts):
            return Instruction("Q", True, m.ref_pos, m.mut_pos, 0, "")
        return _phi_ins()
    return _i_s_frameshift(m, muts)


def _i_inframe_deletion_and_stop_retained(m, muts):  # :869-889
    i = _i_stop_gained(m, muts)
    i.code = "P"
    if m.ref_aa[0] == "End":
        i.len = len(m.ref_aa[1]) - 1
    return i


def _i_stop_lost_and_frameshift(m, muts):  # :967-974
    if m.ref_aa[0] == "Not":
        return _i_stop_lost(m, muts)
    return _i_frameshift(m, muts)


def _i_missense_and_inframe_altering(m, muts):  # :993-1044
    if m.mut_aa[0] == "Not":
        return _recode(_i_frameshift(m, muts), "Y")  # frameshift of NotSeq is phi -> stays 'E'
    if m.ref_aa[0] == "Not":
        raise RefPanic("missense&inframe_altering with '*' reference")
    return _code2or3(m)


def _i_start_lost_and_splice_region(m, muts):  # :1066-1071
    i = _i_start_lost(m, muts)
    i.code = "U"
    return i


_INTERP = {  # instruction.rs:64-91
    "missense": _i_missense,
    "*missense": _i_s_missense,
    "frameshift": _i_frameshift,
    "*frameshift": _i_s_frameshift,
    "inframe_insertion": _i_inframe_insertion,
    "*inframe_insertion": _i_s_inframe_insertion,
    "inframe_deletion": _i_inframe_deletion,
    "*inframe_deletion": _i_s_inframe_deletion,
    "start_lost": _i_start_lost,
    "stop_lost": _i_stop_lost,
    "stop_gained": _i_stop_gained,
    "*stop_gained": _i_s_stop_gained,
    "*missense&inframe_altering": lambda m, v: _recode(_i_s_frameshift(m, v), "K"),  # :738-746
    "*frameshift&stop_retained": _i_s_frameshift_and_stop_retained,
    "*stop_gained&inframe_altering": lambda m, v: _recode(_i_s_stop_gained(m, v), "A"),  # :815-823
    "frameshift&stop_retained": lambda m, v: _recode(_i_frameshift(m, v), "B"),  # :842-850
    "inframe_deletion&stop_retained": _i_inframe_deletion_and_stop_retained,
    "inframe_insertion&stop_retained": lambda m, v: _phi_ins(),  # :908-921 always phi
    "stop_gained&inframe_altering": lambda m, v: _recode(_i_stop_gained(m, v), "T"),  # :940-948
    "stop_lost&frameshift": _i_stop_lost_and_frameshift,
    "missense&inframe_altering": _i_missense_and_inframe_altering,
    "start_lost&splice_region": _i_start_lost_and_splice_region,
}


def instruction_from_mutation(m: Mutation, muts: List[Mutation]) -> Instruction:
    return _INTERP[m.mut_type](m, muts)


# --------------------------------------------------------------------------- per transcript
@dataclass
class TranscriptGIR:
    """One transcript's GIR (gir.rs:15-23): tasks + annotation + alt/ref/res tapes."""
    name: str
    tasks: List[Task]
    annotation: Tuple[int, int]
    alt: str
    ref: str
    res_len: int


@dataclass
class TranscriptInstruction:
    name: str
    ref_len: int
    instructions: List[Instruction]

    # transcript_instructions.rs:41-169
    @staticmethod
    def from_alt_transcript(name: str, muts: List[Mutation], ref_seqs: Dict[str, str],
                            inspect_ins_gen: bool = False) -> "TranscriptInstruction":
        muts = sorted(muts, key=lambda m: m.mut_pos)  # sort_alterations (unstable in the reference)
        if name not in ref_seqs:
            raise TaskGenError("transcript %s not in the reference" % name)
        ref_len = len(ref_seqs[name])
        ins = []
        for m in muts:
            i = instruction_from_mutation(m, muts)
            if i.code != "E":
                ins.append(i)
        if not ins:
            raise TaskGenError("no supported mutation in %s" % name)
        if inspect_ins_gen:  # env INSPECT_INS_GEN (:65-156)
            if len({i.pos_ref for i in ins}) != len(ins):
                raise TaskGenError("two mutations at the same position")
            if len(ins) > 1 and not any(i.code == "0" for i in ins):
                for a, b in zip(ins[:-1], ins[1:]):
                    if b.pos_res <= a.pos_res + len(a.data) - 1:
                        raise TaskGenError("mutations overlap")
                    if a.code in "CD" and b.pos_ref <= a.pos_res + a.len - 1:
                        raise TaskGenError("mutations overlap a deletion")
        return TranscriptInstruction(name, ref_len, ins)

    def alt_stream_size(self) -> int:  # :198-206
        return sum(len(i.data) for i in self.instructions)

    def _no_prior_gf(self, ins: Instruction) -> bool:
        idx = next(k for k, e in enumerate(self.instructions) if e == ins)
        return not any(e.code in "GF" for e in self.instructions[:idx])

    def expected_results_size(self) -> int:  # :214-321
        L = self.ref_len
        d = 0
        for i in self.instructions:
            c = i.code
            tail = len(i.data) - (L - i.pos_ref)
            if c in "U0":
                d -= L
                break
            elif c == "F":
                d += tail
            elif c in "RKQ":
                if self._no_prior_gf(i):
                    d += tail
            elif c in "GXT":
                d -= L - i.pos_ref
            elif c in "MN2":
                pass
            elif c == "L":
                if i.pos_ref + 1 == L or i.pos_ref == L:
                    d += len(i.data)
                else:
                    d += tail
            elif c == "I":
                d += len(i.data) - 1
            elif c == "J":
                if self._no_prior_gf(i):
                    d += len(i.data) - 1
            elif c == "D":
                d -= i.len
            elif c == "C":
                if self._no_prior_gf(i):
                    d -= i.len
            elif c == "A":
                if self._no_prior_gf(i):
                    d -= L - i.pos_ref
            elif c == "B":
                d -= L - i.pos_ref - i.len
            elif c == "P":
                d -= i.len
            elif c == "Z":
                pass
            elif c == "W":
                d += len(i.data)
            elif c == "Y":
                d += tail + 1
            elif c == "3":
                d += len(i.data) - i.len
            else:
                raise RefPanic("instruction %r is not supported" % (i,))
        size = L + d
        if size < 0:
            raise RefPanic("negative result size cast to usize")
        return size

    # transcript_instructions.rs:335-427
    def get_g_rep(self, ref_seqs: Dict[str, str]) -> TranscriptGIR:
        if any(i.code in "0U" for i in self.instructions) or not self.instructions:
            return TranscriptGIR(self.name, [], (0, 0), "", "", 0)
        res_len = self.expected_results_size()
        ref = ref_seqs[self.name]
        alt: List[str] = []  # list of chars
        tasks: List[Task] = [_base_task(self.instructions[0], self.ref_len)]
        for ins in self.instructions:
            t1, t2 = _to_task(ins, self.instructions, alt, tasks, len(ref), self.name)
            if t1[0] != 2:
                tasks.append(t1)
            if t2[0] != 2:
                tasks.append(t2)
        return TranscriptGIR(self.name, tasks, (0, res_len), "".join(alt), ref, res_len)


def _base_task(ins: Instruction, ref_len: int) -> Task:  # :713-736
    if ins.code in "ZY":
        return (0, 0, ins.pos_ref + 1, 0)
    if ins.code == "L":
        if ins.pos_ref + 1 == ref_len:
            return (0, 0, ins.pos_ref + 1, 0)
        if ins.pos_ref == ref_len:
            return (0, 0, ins.pos_ref, 0)
        return (0, 0, ins.pos_res, 0)
    return (0, 0, ins.pos_ref, 0)


def _usub(a: int, b: int) -> int:
    """usize subtraction: underflow panics in the reference."""
    if a < b:
        raise RefPanic("attempt to subtract with overflow")
    return a - b


_LAST_ONLY_WHEN_LAST = "KYQABPZTWGFRLX"  # :486
_LAST_ONLY_GUARD = "KQABPZTWGFRL"  # :496  ('X' and 'Y' are missing from the guard list)


def _to_task(ins: Instruction, all_ins: List[Instruction], alt: List[str], tasks: List[Task],
             ref_len: int, name: str) -> Tuple[Task, Task]:  # :452-505
    c = ins.code
    last = tasks[-1]
    pos_result = last[3] + last[2]
    if c in "MN":  # :654-663  alt tape receives the residue twice
        alt.extend(ins.data)
        alt.extend(ins.data)
        t1 = (1, len(alt) - len(ins.data), 1, pos_result)
    elif c in "FRKBY":  # :666-679
        alt.extend(ins.data)
        t1 = (1, len(alt) - len(ins.data), ins.len, pos_result)
    elif c in "GXAT":  # :682-693
        t1 = PHI
    elif c in "LW":  # :696-710
        alt.extend(ins.data)
        t1 = (1, len(alt) - len(ins.data), len(ins.data), pos_result)
    elif c in "IJ2":  # :739-747, :761-769
        off = len(alt)
        alt.extend(ins.data)
        t1 = (1, off, ins.len, pos_result)
    elif c in "DC3":  # :750-758, :772-780
        off = len(alt)
        alt.extend(ins.data)
        t1 = (1, off, len(ins.data), pos_result)
    elif c in "QZP":  # :471
        t1 = PHI
    else:
        raise RefPanic("Instruction %r is not supported" % (ins,))
    is_last = all_ins[-1] == ins
    if is_last:
        if c in _LAST_ONLY_WHEN_LAST:
            t2 = PHI
        else:
            t2 = _add_last(ref_len, ins, t1[3] + t1[2])
    else:
        if c in _LAST_ONLY_GUARD:
            raise TaskGenError("Translating %s failed: %r must be the last mutation" % (name, ins))
        t2 = _add_till_next(ins, all_ins, t1, ref_len)
    return t1, t2


def _add_till_next(ins: Instruction, all_ins: List[Instruction], last_task: Task, ref_len: int) -> Task:
    """:508-629"""
    position = next(k for k, e in enumerate(all_ins) if e == ins)
    nxt = all_ins[position + 1]
    res = last_task[3] + last_task[2]
    c = ins.code
    if c in "DC":
        if nxt.pos_ref == ins.pos_ref:
            return PHI
        if ins.pos_ref + ins.len == nxt.pos_ref:
            return PHI
        start = ins.pos_ref + ins.len + 1
        if nxt.code == "L" and nxt.pos_ref + 1 == ref_len and start == nxt.pos_ref:
            return (0, start, 1, res)
        return (0, start, _usub(nxt.pos_ref, start), res)
    if c in "23":
        if nxt.pos_ref == ins.pos_ref:
            return PHI
        if ins.pos_ref + ins.len == nxt.pos_ref:
            return PHI
        start = ins.pos_ref + ins.len
        return (0, start, _usub(nxt.pos_ref, start), res)
    if nxt.pos_ref == ins.pos_ref:
        return PHI
    if nxt.code == "L" and nxt.pos_ref + 1 == ref_len:
        return (0, ins.pos_ref + 1, _usub(nxt.pos_ref, ins.pos_ref), res)
    return (0, ins.pos_ref + 1, _usub(_usub(nxt.pos_ref, 1), ins.pos_ref), res)


def _add_last(ref_len: int, ins: Instruction, pos_res: int) -> Task:  # :633-651
    if ins.code in "DC":
        return (0, ins.pos_ref + ins.len + 1, _usub(_usub(_usub(ref_len, ins.pos_ref), ins.len), 1), pos_res)
    if ins.code in "23":
        return (0, ins.pos_ref + ins.len, _usub(_usub(ref_len, ins.pos_ref), ins.len), pos_res)
    return (0, ins.pos_ref + 1, _usub(_usub(ref_len, ins.pos_ref), 1), pos_res)


# --------------------------------------------------------------------------- per haplotype
@dataclass
class HaplotypeGIR:
    """The engine's input for one haplotype (gir.rs:15-23 after haplotype_instruction.rs:75-137)."""
    tasks: List[Task]
    annotation: Dict[str, Tuple[int, int]]  # insertion-ordered; the reference's HashMap order is random
    alt: str
    ref: str
    res_len: int
    skipped: List[str] = field(default_factory=list)


def haplotype_instructions(alt_transcripts: Sequence[Tuple[str, List[Mutation]]],
                           ref_seqs: Dict[str, str]) -> List[TranscriptInstruction]:
    """haplotype_instruction.rs:37-72 -- Err(..) transcripts are filtered out."""
    out = []
    for name, muts in alt_transcripts:
        try:
            out.append(TranscriptInstruction.from_alt_transcript(name, muts, ref_seqs))
        except TaskGenError:
            pass
    return out


def haplotype_g_rep(instrs: List[TranscriptInstruction], ref_seqs: Dict[str, str]) -> HaplotypeGIR:
    """haplotype_instruction.rs:75-158 -- concatenate + re-index.

    The result tape is sized from ALL instructions (:78, :161-168) before any transcript can
    fail in get_g_rep, so a skipped transcript leaves un-annotated trailing '.' residues.
    """
    res_len = sum(t.expected_results_size() for t in instrs)
    tasks: List[Task] = []
    annotation: Dict[str, Tuple[int, int]] = {}
    alt_parts: List[str] = []
    ref_parts: List[str] = []
    skipped: List[str] = []
    ref_c = alt_c = res_c = 0
    for ti in instrs:
        try:
            g = ti.get_g_rep(ref_seqs)
        except TaskGenError:
            skipped.append(ti.name)
            continue
        for (code, sp, ln, spr) in g.tasks:
            if code == 0:
                tasks.append((0, sp + ref_c, ln, spr + res_c))
            elif code == 1:
                tasks.append((1, sp + alt_c, ln, spr + res_c))
            else:
                raise RefPanic("Unsupported Stream code")  # :154
        alt_parts.append(g.alt)
        ref_parts.append(g.ref)
        annotation[g.name] = (g.annotation[0] + res_c, g.annotation[1] + res_c)
        ref_c += len(g.ref)
        alt_c += len(g.alt)
        res_c += g.res_len
    return HaplotypeGIR(tasks, annotation, "".join(alt_parts), "".join(ref_parts), res_len, skipped)


def group_muts_per_transcript(csqs: Sequence[str]) -> List[Tuple[str, List[Mutation]]]:
    """vcf_tools.rs:82-133: unique transcript ids, sorted; membership is a SUBSTRING test (:90)."""
    names = []
    for c in csqs:
        try:
            names.append(split_csq_string(c)[1])
        except TaskGenError:
            pass
    out = []
    for name in sorted(set(names)):
        out.append((name, alt_transcript(name, [c for c in csqs if name in c])))
    return out


# --------------------------------------------------------------------------- pure-python engine + consumer
def execute_tasks(tasks: Sequence[Task], ref: str, alt: str, res_len: int, fill: str = ".") -> str:
    """task.rs:38-50 looped as gir.rs:230-234, on python strings (small cases only)."""
    res = [fill] * res_len
    for (code, sp, ln, spr) in tasks:
        src = ref if code == 0 else alt
        if spr + ln > len(res) or sp + ln > len(src):
            raise RefPanic("slice index out of range")
        res[spr:spr + ln] = src[sp:sp + ln]
    return "".join(res)


def sequence_tape_records(tape: str, annotation: Dict[str, Tuple[int, int]], hap: int) -> List[Tuple[str, str]]:
    """sequence_tape.rs:33-41,77-89 + personalized_genome.rs:72-117: (header, sequence) records."""
    mx = max([e for (_, e) in annotation.values()], default=0)
    if mx > len(tape):
        raise RefPanic("Bad Tape Encountered")
    return [("%s_%d" % (k, hap), tape[s:e]) for k, (s, e) in annotation.items()]


def haplotype_records(csqs: Sequence[str], ref_seqs: Dict[str, str], hap: int) -> List[Tuple[str, str]]:
    """csq strings of one haplotype -> FASTA records, end to end through the restatement."""
    g = haplotype_g_rep(haplotype_instructions(group_muts_per_transcript(csqs), ref_seqs), ref_seqs)
    tape = execute_tasks(g.tasks, g.ref, g.alt, g.res_len)
    return sequence_tape_records(tape, g.annotation, hap)
