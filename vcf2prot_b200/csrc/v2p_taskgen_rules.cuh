// v2p_taskgen_rules.cuh -- the reference's Instruction -> Task emission rules for ONE transcript on one haplotype, for every
// instruction code instruction.rs can produce.  Plain functions (host + device): the CUDA generator (v2p_taskgen.cu) runs
// them one thread per transcript-on-haplotype, and tests/cpp/taskgen_rules_test.cpp compiles the same text with g++ to
// hold it to the reference's own unit-test vectors without a GPU.
//
// Restated (paths under /root/reference/src/data_structures/InternalRep/):
//   instruction.rs:1075-1098            validate_s_state  ('*'-prefixed consequence classes are dropped when an earlier
//                                       mutation of the transcript already ended or shifted the protein)
//   transcript_instructions.rs:41-63    from_alt_transcript: phi ('E') instructions are dropped; none left -> Err
//   transcript_instructions.rs:214-321  compute_expected_results_array_size  (per-code size deltas)
//   transcript_instructions.rs:335-427  get_g_rep: start_lost ('0'/'U') -> empty GIR; base task; per-instruction tasks
//   transcript_instructions.rs:452-505  to_task, :508-629 add_till_next_ins, :633-651 add_last_instruction,
//                                       :654-780 the per-code task builders, :713-736 build_base_instruction
// A Task is (stream, src, len, dst): src is transcript-relative (stream 0) or relative to the transcript's own
// alteration bytes (stream 1); dst is relative to the transcript's result.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define V2P_HD __host__ __device__ __forceinline__
#else
#define V2P_HD inline
#endif

namespace v2p_rules {

enum : uint8_t { INS_STAR = 1, INS_INVALIDATES = 2 };  // TgIns::flags
// STAR        the csq class is '*'-prefixed: Instruction::from_mutation calls validate_s_state (instruction.rs:222-238 ...)
// INVALIDATES mut_type is stop_gained / frameshift / *stop_gained, or inframe_insertion / inframe_deletion whose mutated
//             amino-acid field is '*' or ends in '*' (instruction.rs:1083-1094)

struct TgIns {  // instruction.rs:6-16 with the data held by (offset, length) into a pool
    uint8_t code;   // 'M','N','I','J','D','C','2','3','F','R','K','B','Y','L','W','G','X','A','T','P','Q','Z','0','U'; 'E' = phi
    uint8_t flags;  // INS_*
    uint32_t pos_ref, pos_res, len, dlen;
    uint64_t doff;
};

enum {
    TG_OK = 0,
    TG_EMPTY = 1,    // start_lost: empty GIR, annotation (s, s)                       transcript_instructions.rs:338-343
    TG_ABSENT = 2,   // no supported mutation left: the transcript is not in the haplotype at all      :60-63 (Err)
    TG_SKIPPED = 3,  // "... must be the last mutation": get_g_rep fails, the transcript is skipped but its expected size
                     // was already counted into the haplotype's tape (trailing '.')              :496-499, hap_ins:78
    TG_PANIC = 4     // the reference aborts here (usize underflow, negative size)
};

V2P_HD bool in_set(uint8_t c, const char* set) {
    for (; *set; ++set)
        if ((uint8_t)*set == c) return true;
    return false;
}

struct TgSummary {
    int status;
    uint64_t size;     // expected_results_size (counted into the haplotype tape unless TG_ABSENT)
    uint32_t n_tasks;  // tasks pushed (0 unless TG_OK)
    uint64_t n_alt;    // alteration bytes pushed (0 unless TG_OK)
};

// Sink concept:  void task(uint32_t stream, uint64_t src, uint64_t len, uint64_t dst);
//                void alt(uint64_t doff, uint32_t dlen);       // append pool[doff .. doff+dlen) to the transcript's alt bytes
struct NullSink {
    V2P_HD void task(uint32_t, uint64_t, uint64_t, uint64_t) {}
    V2P_HD void alt(uint64_t, uint32_t) {}
};

// get(i) -> TgIns of the i-th mutation of the transcript on this haplotype (sorted by mutated position), i in [0, n).
template <class Get, class Sink>
V2P_HD TgSummary tg_transcript(const Get& get, const int n, const uint64_t ref_len, Sink& sink) {
    TgSummary s{TG_OK, 0, 0, 0};
    // ---- which instructions survive: phi dropped, '*' classes dropped after an invalidating mutation
    int n_kept = 0, last_kept = -1;
    bool has_start_lost = false;
    {
        bool inval = false;
        for (int i = 0; i < n; ++i) {
            const TgIns it = get(i);
            const bool kept = it.code != 'E' && !((it.flags & INS_STAR) && inval);
            inval = inval || (it.flags & INS_INVALIDATES);
            if (kept) {
                ++n_kept;
                last_kept = i;
                has_start_lost = has_start_lost || it.code == '0' || it.code == 'U';
            }
        }
    }
    if (n_kept == 0) {
        s.status = TG_ABSENT;
        return s;
    }
    auto kept_at = [&](int i, bool inval_before) {
        const TgIns it = get(i);
        return it.code != 'E' && !((it.flags & INS_STAR) && inval_before);
    };
    // ---- expected size of the result (transcript_instructions.rs:214-321)
    {
        long long d = 0;
        bool inval = false, prior_gf = false;
        for (int i = 0; i < n; ++i) {
            const TgIns it = get(i);
            const bool kept = kept_at(i, inval);
            inval = inval || (it.flags & INS_INVALIDATES);
            if (!kept) continue;
            const uint8_t c = it.code;
            const long long L = (long long)ref_len, tail = (long long)it.dlen - (L - (long long)it.pos_ref);
            if (c == 'U' || c == '0') {
                d -= L;
                break;
            } else if (c == 'F') d += tail;
            else if (in_set(c, "RKQ")) d += prior_gf ? 0 : tail;
            else if (in_set(c, "GXT")) d -= L - (long long)it.pos_ref;
            else if (in_set(c, "MN2")) {}
            else if (c == 'L') d += ((uint64_t)it.pos_ref + 1 == ref_len || it.pos_ref == ref_len) ? (long long)it.dlen : tail;
            else if (c == 'I') d += (long long)it.dlen - 1;
            else if (c == 'J') d += prior_gf ? 0 : (long long)it.dlen - 1;
            else if (c == 'D') d -= it.len;
            else if (c == 'C') d -= prior_gf ? 0 : (long long)it.len;
            else if (c == 'A') d -= prior_gf ? 0 : L - (long long)it.pos_ref;
            else if (c == 'B') d -= L - (long long)it.pos_ref - (long long)it.len;
            else if (c == 'P') d -= it.len;
            else if (c == 'Z') {}
            else if (c == 'W') d += it.dlen;
            else if (c == 'Y') d += tail + 1;
            else if (c == '3') d += (long long)it.dlen - (long long)it.len;
            else {
                s.status = TG_PANIC;  // "instruction ... is not supported"
                return s;
            }
            prior_gf = prior_gf || c == 'G' || c == 'F';
        }
        if ((long long)ref_len + d < 0) {
            s.status = TG_PANIC;
            return s;
        }
        s.size = (uint64_t)((long long)ref_len + d);
    }
    if (has_start_lost) {
        s.status = TG_EMPTY;
        return s;
    }
    // ---- tasks (get_g_rep :345-427): base task, then (t1, t2) per instruction; phi tasks are not pushed, and every
    //      destination is "end of the last PUSHED task" (:657 etc.)
    uint64_t last_dst = 0, last_len = 0, alt_len = 0;
    uint32_t n_tasks = 0;
    bool panic = false;
    auto usub = [&](uint64_t a, uint64_t b) -> uint64_t {  // usize subtraction: underflow aborts the reference
        if (a < b) {
            panic = true;
            return 0;
        }
        return a - b;
    };
    auto push = [&](uint32_t stream, uint64_t src, uint64_t len, uint64_t dst) {
        sink.task(stream, src, len, dst);
        last_dst = dst, last_len = len;
        ++n_tasks;
    };
    bool first = true, inval = false;
    for (int i = 0; i < n; ++i) {
        const TgIns it = get(i);
        const bool kept = kept_at(i, inval);
        inval = inval || (it.flags & INS_INVALIDATES);
        if (!kept) continue;
        const uint8_t c = it.code;
        if (first) {  // build_base_instruction :713-736
            uint64_t bl = it.pos_ref;
            if (c == 'Z' || c == 'Y') bl = (uint64_t)it.pos_ref + 1;
            else if (c == 'L') bl = ((uint64_t)it.pos_ref + 1 == ref_len) ? (uint64_t)it.pos_ref + 1 : (it.pos_ref == ref_len ? it.pos_ref : it.pos_res);
            push(0, 0, bl, 0);
            first = false;
        }
        const bool is_last = i == last_kept;
        // t1: the mutation itself
        bool t1_phi = false;
        uint64_t t1_dst = last_dst + last_len, t1_len = 0;
        if (c == 'M' || c == 'N') {  // :654-663: the alt tape receives the residue(s) twice; the task reads the second copy
            sink.alt(it.doff, it.dlen);
            sink.alt(it.doff, it.dlen);
            alt_len += 2ull * it.dlen;
            t1_len = 1;
            push(1, alt_len - it.dlen, 1, t1_dst);
        } else if (in_set(c, "FRKBY")) {  // :666-679
            sink.alt(it.doff, it.dlen);
            alt_len += it.dlen;
            t1_len = it.len;
            push(1, alt_len - it.dlen, it.len, t1_dst);
        } else if (in_set(c, "GXAT") || in_set(c, "QZP")) {  // :682-693, :471
            t1_phi = true;
        } else if (c == 'L' || c == 'W') {  // :696-710
            sink.alt(it.doff, it.dlen);
            alt_len += it.dlen;
            t1_len = it.dlen;
            push(1, alt_len - it.dlen, it.dlen, t1_dst);
        } else if (in_set(c, "IJ2")) {  // :739-747, :761-769
            const uint64_t off = alt_len;
            sink.alt(it.doff, it.dlen);
            alt_len += it.dlen;
            t1_len = it.len;
            push(1, off, it.len, t1_dst);
        } else if (in_set(c, "DC3")) {  // :750-758, :772-780
            const uint64_t off = alt_len;
            sink.alt(it.doff, it.dlen);
            alt_len += it.dlen;
            t1_len = it.dlen;
            push(1, off, it.dlen, t1_dst);
        } else {
            s.status = TG_PANIC;
            return s;
        }
        // t2: the reference copy that follows.  Its destination is t1's end -- and t1 = phi is the tuple (2,0,0,0), whose
        // "end" is 0 (:503, :511): a follow-up copy after a phi mutation task lands at result offset 0.
        const uint64_t t2_dst = t1_phi ? 0 : t1_dst + t1_len;
        if (is_last) {
            if (in_set(c, "KYQABPZTWGFRLX")) continue;  // :486
            uint64_t start, len;  // add_last_instruction :633-651
            if (c == 'D' || c == 'C') start = (uint64_t)it.pos_ref + it.len + 1, len = usub(usub(usub(ref_len, it.pos_ref), it.len), 1);
            else if (c == '2' || c == '3') start = (uint64_t)it.pos_ref + it.len, len = usub(usub(ref_len, it.pos_ref), it.len);
            else start = (uint64_t)it.pos_ref + 1, len = usub(usub(ref_len, it.pos_ref), 1);
            if (panic) break;
            push(0, start, len, t2_dst);
        } else {
            if (in_set(c, "KQABPZTWGFRL")) {  // :496-499 ('X' and 'Y' are missing from the reference's list)
                s.status = TG_SKIPPED;
                return s;
            }
            TgIns nx{};  // the next surviving instruction
            {
                bool iv = inval;
                for (int k = i + 1; k < n; ++k) {
                    const TgIns cand = get(k);
                    const bool kk = cand.code != 'E' && !((cand.flags & INS_STAR) && iv);
                    iv = iv || (cand.flags & INS_INVALIDATES);
                    if (kk) {
                        nx = cand;
                        break;
                    }
                }
            }
            // add_till_next_ins :508-629
            if (nx.pos_ref == it.pos_ref) continue;
            if (c == 'D' || c == 'C') {
                if ((uint64_t)it.pos_ref + it.len == nx.pos_ref) continue;
                const uint64_t start = (uint64_t)it.pos_ref + it.len + 1;
                if (nx.code == 'L' && (uint64_t)nx.pos_ref + 1 == ref_len && start == nx.pos_ref) push(0, start, 1, t2_dst);
                else {
                    const uint64_t len = usub(nx.pos_ref, start);
                    if (panic) break;
                    push(0, start, len, t2_dst);
                }
            } else if (c == '2' || c == '3') {
                if ((uint64_t)it.pos_ref + it.len == nx.pos_ref) continue;
                const uint64_t start = (uint64_t)it.pos_ref + it.len;
                const uint64_t len = usub(nx.pos_ref, start);
                if (panic) break;
                push(0, start, len, t2_dst);
            } else if (nx.code == 'L' && (uint64_t)nx.pos_ref + 1 == ref_len) {
                const uint64_t len = usub(nx.pos_ref, it.pos_ref);
                if (panic) break;
                push(0, (uint64_t)it.pos_ref + 1, len, t2_dst);
            } else {
                const uint64_t len = usub(usub(nx.pos_ref, 1), it.pos_ref);
                if (panic) break;
                push(0, (uint64_t)it.pos_ref + 1, len, t2_dst);
            }
        }
    }
    if (panic) {
        s.status = TG_PANIC;
        return s;
    }
    s.n_tasks = n_tasks;
    s.n_alt = alt_len;
    return s;
}

}  // namespace v2p_rules
