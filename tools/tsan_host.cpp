// ThreadSanitizer harness for the threaded host code (piece-wise P-line tokenisation, block-wise table
// formatting): `make tsan && FGFA_GPU_PARSE=0 build/tsan/tsan_host big.gfa` -- expects no TSAN report.
#include <cstdio>
#include <string>
#include <vector>
#include "../pollen_b200/csrc/file.hpp"
#include "../pollen_b200/csrc/ops_depth.hpp"
#include "../pollen_b200/csrc/parse.hpp"
#include "../pollen_b200/csrc/print.hpp"
int main(int argc, char** argv) {
    flatgfa::MappedFile f(argv[1]);
    flatgfa::HeapGFAStore s = flatgfa::Parser::parse_mem(f.data(), f.size());   // threaded pieces
    flatgfa::FlatGFA g = s.view();
    std::string text;
    flatgfa::print::gfa(g, text);
    std::vector<uint64_t> d(g.segs.len(), 12345678901ull), u(g.segs.len(), 7);
    size_t len = 0;
    char* t = flatgfa::ops::depth::seg_depth_table(g, d.data(), u.data(), &len);   // threaded blocks
    std::printf("steps=%zu text=%zu table=%zu\n", g.steps.len(), text.size(), len);
    free(t);
    return 0;
}
