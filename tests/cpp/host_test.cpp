// C++ host-side tests of include/v2p_host.hpp, written after the reference's own unit tests:
//   task.rs:118-144 test_execute, engines.rs doc-test, transcript_instructions.rs:884-1594 test_correct_translation_N
// (cases arrive as a text fixture exported from tests/golden by tests/test_cpp_host.py).
//   host_test --no-gpu            selector + batch re-indexing only (no device call)
//   host_test cases.txt           everything, on cuda:0
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "v2p_host.hpp"

static int failures = 0;
#define CHECK(cond)                                                            \
    do {                                                                       \
        if (!(cond)) {                                                         \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);        \
            ++failures;                                                        \
        }                                                                      \
    } while (0)

static std::u32string widen(const std::string& s) { return std::u32string(s.begin(), s.end()); }
static std::string narrow(const std::u32string& s) { return std::string(s.begin(), s.end()); }

static void test_engine_from_str() {  // engines.rs:17-30
    CHECK(v2p::engine_from_str("st") == v2p::Engine::ST && v2p::engine_from_str("ST") == v2p::Engine::ST);
    CHECK(v2p::engine_from_str("mt") == v2p::Engine::MT && v2p::engine_from_str("MT") == v2p::Engine::MT);
    CHECK(v2p::engine_from_str("gpu") == v2p::Engine::GPU && v2p::engine_from_str("GPU") == v2p::Engine::GPU);
    for (const char* bad : {"Gpu", "cuda", "", "st "}) {
        bool threw = false;
        try { v2p::engine_from_str(bad); } catch (const v2p::EngineError& e) { threw = e.status == V2P_ERR_BAD_ENGINE; }
        CHECK(threw);
    }
}

static void test_batch_reindexing() {  // haplotype_instruction.rs:94-158 (golden: SURVEY 8c "concat" row)
    const std::string R = "MEDLGENTMVLSTLRSLNNFISQRVEGGSGLEELERGG";
    v2p::HaplotypeBatch b;
    b.begin_haplotype();
    b.add_transcript("TA", {v2p::Task(0, 0, 37, 0)}, "", R, 0, 38);                                                   // tr_20: '.' gap
    b.add_transcript("TB", {v2p::Task(0, 0, 4, 0), v2p::Task(1, 1, 1, 4), v2p::Task(0, 5, 33, 5)}, "HH", R, 38, 38);  // tr_1
    const auto& t = b.tasks();
    CHECK(t.size() == 4);
    // the reference's haplotype table: (0,0,37,0)(0,38,4,38)(1,1,1,42)(0,43,33,43)
    CHECK(t[1].src_off == 38 && t[1].len == 4 && t[1].dst_off == 38 && t[1].stream == 0);
    CHECK(t[2].src_off == 1 && t[2].len == 1 && t[2].dst_off == 42 && t[2].stream == 1);
    CHECK(t[3].src_off == 43 && t[3].len == 33 && t[3].dst_off == 43);
    CHECK(b.annotation(0).at("TB") == std::make_pair(uint64_t(38), uint64_t(76)));
    bool threw = false;
    try { b.add_transcript("TC", {v2p::Task(2, 0, 0, 0)}, "", R, 76, 38); } catch (const v2p::EngineError& e) { threw = e.status == V2P_ERR_BAD_STREAM; }
    CHECK(threw);  // haplotype_instruction.rs:154
}

static void test_sequence_tape() {  // sequence_tape.rs:17-31 doc-test, :95-104, :122-170 (bad tape)
    const std::string code = "SEQ1_SEQ2_SEQ3_SEQ4_SEQ5_SEQ6";
    v2p::Annotation m{{"1", {0, 4}}, {"2", {5, 9}}, {"3", {10, 14}}};
    v2p::SequenceTape t(code, m);
    CHECK(t.get_seq("1") == "SEQ1" && t.get_seq("2") == "SEQ2" && t.get_seq("3") == "SEQ3");
    v2p::Annotation six{{"1", {0, 4}}, {"2", {5, 9}}, {"3", {10, 14}}, {"4", {15, 19}}, {"5", {20, 24}}, {"6", {25, 29}}};
    CHECK(v2p::SequenceTape::get_max_index(six) == 29);
    bool threw = false;
    try { v2p::SequenceTape bad("SEQ1", v2p::Annotation{{"1", {0, 5}}}); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    threw = false;
    try { t.get_seq("9"); } catch (const std::out_of_range&) { threw = true; }
    CHECK(threw);
    CHECK(v2p::SequenceTape("MK", v2p::Annotation{{"T", {0, 2}}, {"U", {2, 2}}}).fasta_text(2) == ">T_2\nMK\n>U_2\n\n");  // empty record
}

static void test_task_rs_vector(v2p::Context& ctx) {  // task.rs:118-144
    v2p::GIR g({v2p::Task(0, 1, 1, 8), v2p::Task(0, 4, 1, 4), v2p::Task(0, 6, 2, 6)}, {}, widen("HGFEFCBA"), widen("ABCFEFGH"),
               widen("xxxxxxxxxx"));
    auto res = std::move(g).execute(v2p::Engine::GPU, ctx);
    CHECK(narrow(res.first) == "xxxxExGHBx");
    bool threw = false;
    try {
        v2p::GIR g2({v2p::Task(0, 0, 1, 0)}, {}, widen(""), widen("A"), widen("."));
        std::move(g2).execute(v2p::Engine::ST, ctx);
    } catch (const v2p::EngineError& e) { threw = e.status == V2P_ERR_NOT_GPU_ENGINE; }
    CHECK(threw);
    threw = false;
    try {  // task.rs:44 slice panic
        v2p::GIR g3({v2p::Task(0, 0, 9, 0)}, {}, widen("xyz"), widen("ABCDEFGH"), widen("........"));
        std::move(g3).execute(v2p::Engine::GPU, ctx);
    } catch (const v2p::EngineError& e) { threw = e.status == V2P_ERR_RES_OOB && e.bad_task == 0; }
    CHECK(threw);
}

struct Case {
    std::string name, ref, alt, expect;
    uint64_t res_len = 0;
    std::vector<v2p::Task> tasks;
};

static std::vector<Case> load_cases(const char* path) {
    std::vector<Case> out;
    std::ifstream f(path);
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream is(line);
        std::string key;
        is >> key;
        if (key == "CASE") { out.emplace_back(); is >> out.back().name; }
        else if (key == "REF") { is >> out.back().ref; }
        else if (key == "ALT") { is >> out.back().alt; if (out.back().alt == "-") out.back().alt.clear(); }
        else if (key == "RES") { is >> out.back().res_len; }
        else if (key == "TASK") { uint64_t c, a, l, d; is >> c >> a >> l >> d; out.back().tasks.emplace_back((uint8_t)c, a, l, d); }
        else if (key == "EXPECT") { is >> out.back().expect; if (out.back().expect == "-") out.back().expect.clear(); }
    }
    return out;
}

// The reference's own unit-test mutations (transcript_instructions.rs:884, :987, :1053, :1306) as two probands, from
// Instruction values to {dir}/{proband}.fasta through the device generator, the engine and the native writer.
static std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    return std::string((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static void test_pipeline_writes_the_golden_records(v2p::Context& ctx, const std::string& dir) {
    const std::string T = "MEDLGENTMVLSTLRSLNNFISQRVEGGSGLEELERGG";
    ctx.set_reference(T);
    // sorted by (transcript, mutated position): I and N share position 4 on different haplotypes
    const std::vector<v2p::Instruction> ins = {
        {0, 'I', false, false, 4, 4, 5, "GTEST"},             // inframe_insertion 5G>5GTEST        (test 5)
        {0, 'N', true, false, 4, 4, 1, "H"},                  // *missense 5G>5H                    (test 1)
        {0, 'F', false, true, 9, 9, 15, "VTESTFRAMESHIFT"},   // frameshift 10V>10VTESTFRAMESHIFT   (test 8)
        {0, 'P', false, false, 37, 37, 0, ""},                // inframe_deletion&stop_retained 38*>38*  (test 20)
    };
    v2p::InstructionCatalogue lane0(0, {0, T.size()}, {"T"}, ins), lane1(0, {0, T.size()}, {"T"}, ins);
    v2p::Pipeline pipe(ctx, {&lane0, &lane1});
    v2p::DirWriter w(dir, {"HG1", "HG2", "HG3"}, false, 2);
    const v2p_pipeline_result r = pipe.write({{1}, {2}, {0}, {3}, {}, {}}, w, false, 1);
    CHECK(r.n_samples == 3 && r.n_chunks == 3 && r.n_records == 4 && w.files_written() == 3);
    CHECK(slurp(dir + "/HG1.fasta") == ">T_1\nMEDLHENTMVLSTLRSLNNFISQRVEGGSGLEELERGG\n>T_2\nMEDLGENTMVTESTFRAMESHIFT\n");
    CHECK(slurp(dir + "/HG2.fasta") == ">T_1\nMEDLGTESTENTMVLSTLRSLNNFISQRVEGGSGLEELERGG\n>T_2\nMEDLGENTMVLSTLRSLNNFISQRVEGGSGLEELERG.\n");
    CHECK(slurp(dir + "/HG3.fasta").empty());
    // `-a` (write_all, personalized_genome.rs:120-210): a haplotype without an altered form of T gets the reference record
    pipe.enable_write_all(T, {0, T.size()}, {"T"});
    v2p::DirWriter wa(dir, {"HG1", "HG2", "HG3"}, false, 2);
    const v2p_pipeline_result ra = pipe.write({{1}, {2}, {0}, {3}, {}, {}}, wa, false, 2, /*write_all=*/true);
    CHECK(ra.n_records == 6 && wa.files_written() == 3);
    CHECK(slurp(dir + "/HG1.fasta") == ">T_1\nMEDLHENTMVLSTLRSLNNFISQRVEGGSGLEELERGG\n>T_2\nMEDLGENTMVTESTFRAMESHIFT\n");
    CHECK(slurp(dir + "/HG3.fasta") == ">T_1\n" + T + "\n>T_2\n" + T + "\n");
    // parts/exec.rs:34-40 from one process over several workers (the same GPU twice here): same files
    {
        v2p::Cohort cohort({0, 0}, T, {0, T.size()}, {"T"}, ins, 1);
        v2p::DirWriter wc(dir, {"HG1", "HG2", "HG3"}, false, 2);
        const v2p_cohort_result rc = cohort.write({{1}, {2}, {0}, {3}, {}, {}}, wc, false, 1);
        CHECK(rc.n_devices == 2 && rc.total.n_samples == 3 && rc.total.n_records == 4 && wc.files_written() == 3);
        CHECK(rc.first_sample[0] == 0 && rc.first_sample[2] == 3);
        CHECK(slurp(dir + "/HG1.fasta") == ">T_1\nMEDLHENTMVLSTLRSLNNFISQRVEGGSGLEELERGG\n>T_2\nMEDLGENTMVTESTFRAMESHIFT\n");
        CHECK(slurp(dir + "/HG2.fasta") == ">T_1\nMEDLGTESTENTMVLSTLRSLNNFISQRVEGGSGLEELERGG\n>T_2\nMEDLGENTMVLSTLRSLNNFISQRVEGGSGLEELERG.\n");
        CHECK(slurp(dir + "/HG3.fasta").empty());
        // ... and `-a` over the workers
        cohort.enable_write_all(T, {0, T.size()}, {"T"});
        v2p::DirWriter wd(dir, {"HG1", "HG2", "HG3"}, false, 2);
        const v2p_cohort_result rd = cohort.write({{1}, {2}, {0}, {3}, {}, {}}, wd, false, 1, /*write_all=*/true);
        CHECK(rd.total.n_records == 6 && wd.files_written() == 3);
        CHECK(slurp(dir + "/HG2.fasta") == ">T_1\nMEDLGTESTENTMVLSTLRSLNNFISQRVEGGSGLEELERGG\n>T_2\nMEDLGENTMVLSTLRSLNNFISQRVEGGSGLEELERG.\n");
        CHECK(slurp(dir + "/HG3.fasta") == ">T_1\n" + T + "\n>T_2\n" + T + "\n");
    }
}

int main(int argc, char** argv) {
    test_engine_from_str();
    test_batch_reindexing();
    test_sequence_tape();
    if (argc > 1 && !std::strcmp(argv[1], "--no-gpu")) {
        std::printf("%s (host-only part)\n", failures ? "FAILED" : "OK");
        return failures ? 1 : 0;
    }
    v2p::Context ctx(0);
    test_task_rs_vector(ctx);
    std::vector<Case> cases = argc > 1 ? load_cases(argv[1]) : std::vector<Case>();
    // (1) each golden transcript through GIR::execute(Engine::GPU), like test_correct_translation_N
    for (const Case& c : cases) {
        v2p::GIR g(c.tasks, {{c.name, {0, c.res_len}}}, widen(c.alt), widen(c.ref), std::u32string(c.res_len, U'.'));
        auto res = std::move(g).execute(v2p::engine_from_str("gpu"), ctx);
        CHECK(narrow(res.first) == c.expect);
        CHECK(res.second.at(c.name).second == c.res_len);
    }
    // (2) all of them as ONE haplotype batch, reference layout and shared-proteome layouts
    for (int mode = 0; mode < 3; ++mode) {
        const bool shared = mode > 0, aligned = mode == 2;
        v2p::HaplotypeBatch hb(shared ? v2p::HaplotypeBatch::RefLayout::SharedProteome : v2p::HaplotypeBatch::RefLayout::PerHaplotype,
                               aligned);
        std::string proteome;
        std::vector<uint64_t> offs;
        for (const Case& c : cases) { offs.push_back(proteome.size()); proteome += c.ref; }
        if (shared) ctx.set_reference(proteome);
        for (int h = 0; h < 2; ++h) {
            hb.begin_haplotype();
            for (size_t i = h; i < cases.size(); i += 2)
                hb.add_transcript(cases[i].name + "#" + std::to_string(i), cases[i].tasks, cases[i].alt, cases[i].ref, offs[i],
                                  cases[i].res_len);
        }
        const std::string out = hb.execute(v2p::Engine::GPU, ctx);
        for (int h = 0; h < 2; ++h)
            for (size_t i = h; i < cases.size(); i += 2) {
                auto se = hb.annotation(h).at(cases[i].name + "#" + std::to_string(i));
                CHECK(out.substr(hb.out_base(h) + se.first, se.second - se.first) == cases[i].expect);  // SequenceTape::get_seq
            }
    }
    if (argc > 2) test_pipeline_writes_the_golden_records(ctx, argv[2]);
    std::printf("%s (%zu golden cases x 4 paths)\n", failures ? "FAILED" : "OK", cases.size());
    return failures ? 1 : 0;
}
