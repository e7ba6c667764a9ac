// Host build of the device task generator's rules (vcf2prot_b200/csrc/v2p_taskgen_rules.cuh): reads transcripts as lists
// of instructions, prints what tg_transcript emits.  tests/test_taskgen_rules.py compares that with the oracle
// (reference-pinned restatement of transcript_instructions.rs) -- the same text runs one thread per transcript on the GPU.
//   CASE <name> <ref_len> <n>   then n lines   INS <code> <flags> <pos_ref> <pos_res> <len> <data|->
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "v2p_taskgen_rules.cuh"

using namespace v2p_rules;

struct PrintSink {
    const std::string* pool;
    std::string alt_bytes;
    std::ostringstream out;
    void task(uint32_t stream, uint64_t src, uint64_t len, uint64_t dst) { out << "TASK " << stream << ' ' << src << ' ' << len << ' ' << dst << '\n'; }
    void alt(uint64_t doff, uint32_t dlen) { alt_bytes += pool->substr(doff, dlen); }
};

int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream hs(line);
        std::string tag, name;
        uint64_t ref_len;
        int n;
        if (!(hs >> tag >> name >> ref_len >> n) || tag != "CASE") continue;
        std::vector<TgIns> ins;
        std::string pool;
        for (int i = 0; i < n; ++i) {
            std::getline(std::cin, line);
            std::istringstream is(line);
            std::string t, code, data;
            unsigned flags;
            TgIns x{};
            is >> t >> code >> flags >> x.pos_ref >> x.pos_res >> x.len >> data;
            if (data == "-") data.clear();
            x.code = (uint8_t)code[0], x.flags = (uint8_t)flags, x.doff = pool.size(), x.dlen = (uint32_t)data.size();
            pool += data;
            ins.push_back(x);
        }
        PrintSink sink;
        sink.pool = &pool;
        auto get = [&](int i) { return ins[i]; };
        NullSink null;
        const TgSummary c = tg_transcript(get, n, ref_len, null);  // the counting pass the generator runs first
        const TgSummary s = tg_transcript(get, n, ref_len, sink);
        if (c.status != s.status || c.size != s.size || c.n_tasks != s.n_tasks || c.n_alt != s.n_alt) std::printf("MISMATCH count/emit\n");
        std::printf("RESULT %s %d %llu %u %llu\n", name.c_str(), s.status, (unsigned long long)s.size, s.n_tasks, (unsigned long long)s.n_alt);
        if (s.status == TG_OK) std::printf("%sALT %s\n", sink.out.str().c_str(), sink.alt_bytes.empty() ? "-" : sink.alt_bytes.c_str());
    }
    return 0;
}
