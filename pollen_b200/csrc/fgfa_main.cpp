// `fgfa` command-line tool, restricted to what the node-depth path needs.
//
// Mirrors the reference CLI's spelling (flatgfa/src/cli/main.rs:7-55 global options,
// flatgfa/src/cli/cmds.rs:217-232 `depth` options, main.rs:87-138 input loading and
// dispatch, main.rs:190-213 `dump`):
//   fgfa [-i FLATGFA | -I GFA | < GFA] depth -d          node-depth table on stdout
//   fgfa --gpus N [...] depth -d                         the same table, computed on N GPUs of this box
//                                                        (an addition: the reference has no devices)
//   fgfa [-i FLATGFA | -I GFA | < GFA] depth [-r PATH]   path-depth table (cmds.rs:256-283)
//   fgfa [-i FLATGFA | -I GFA | < GFA] depth -b BED      interval depth table (cmds.rs:246-255)
//   fgfa [-i FLATGFA | -I GFA | < GFA] window-depth PATH SIZE   (cmds.rs:477-496)
//   fgfa [-i FLATGFA | -I GFA | < GFA] -o OUT.flatgfa    convert to the binary format
//   fgfa [-i FLATGFA | -I GFA | < GFA] [-O OUT.gfa]      GFA text to a file or stdout (print.rs)
//   fgfa -m [-p N] -o OUT.flatgfa [-I GFA | < GFA]       preallocated translation (main.rs:216-248)
// The other subcommands of the reference are outside this repository's scope
// and are rejected with an error.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "file.hpp"
#include "ops_depth.hpp"
#include "ops_window_depth.hpp"
#include "parse.hpp"
#include "print.hpp"

namespace {

struct Args {
    std::string input, input_gfa, output, output_gfa;   // -i -I -o -O
    bool mutate = false;                                // -m
    size_t prealloc_factor = 32;                        // -p
    int gpus = 1;                                       // --gpus (not in the reference)
    std::string command;
    // depth (cmds.rs:217-232)
    bool seg_depth = false;                             // -d / --graph-depth-table
    std::vector<std::string> paths;                     // -r
    std::string bed;                                    // -b / --bed-input
    // window-depth (cmds.rs:477-486)
    std::vector<std::string> positional;
};

int usage(const char* msg) {
    if (msg) std::fprintf(stderr, "%s\n", msg);
    std::fprintf(stderr,
                 "Usage: fgfa [-i <input>] [-I <input-gfa>] [-o <output>] [-O <output-gfa>] [-m] "
                 "[-p <prealloc-factor>] [--gpus <n>] [<command>] [<args>]\n\n"
                 "Convert between GFA text and FlatGFA binary formats.\n\n"
                 "Commands:\n  depth             compute depth: the number of times paths cross a node\n"
                 "                    -d, --graph-depth-table  compute node depth instead of path depth\n"
                 "                    -r <path>                in path mode, show only the named path\n"
                 "                    -b, --bed-input <bed>    show depth for intervals from a BED file\n"
                 "  window-depth      find the depth of windows along a path: window-depth <path> <window>\n");
    return 1;
}

bool take(int& i, int argc, char** argv, std::string& dst) {
    if (i + 1 >= argc) return false;
    dst = argv[++i];
    return true;
}

}  // namespace

int main(int argc, char** argv) {
    Args a;
    int i = 1;
    for (; i < argc; ++i) {   // global options, up to the subcommand (main.rs:7-36)
        const std::string t = argv[i];
        if (t == "-i") { if (!take(i, argc, argv, a.input)) return usage("No value provided for option '-i'."); }
        else if (t == "-I") { if (!take(i, argc, argv, a.input_gfa)) return usage("No value provided for option '-I'."); }
        else if (t == "-o") { if (!take(i, argc, argv, a.output)) return usage("No value provided for option '-o'."); }
        else if (t == "-O") { if (!take(i, argc, argv, a.output_gfa)) return usage("No value provided for option '-O'."); }
        else if (t == "-m") a.mutate = true;
        else if (t == "-p") {
            std::string v;
            if (!take(i, argc, argv, v)) return usage("No value provided for option '-p'.");
            char* endp = nullptr;
            const unsigned long long f = std::strtoull(v.c_str(), &endp, 10);
            if (v.empty() || *endp || v[0] == '-') return usage("Error parsing option '-p': not a number");
            a.prealloc_factor = (size_t)f;
        }
        else if (t == "--gpus") {
            std::string v;
            if (!take(i, argc, argv, v)) return usage("No value provided for option '--gpus'.");
            char* endp = nullptr;
            const long g = std::strtol(v.c_str(), &endp, 10);
            if (v.empty() || *endp || g < 1 || g > 16) return usage("Error parsing option '--gpus': expected 1..16");
            a.gpus = (int)g;
        }
        else if (t == "--help" || t == "help") { usage(nullptr); return 0; }
        else if (!t.empty() && t[0] == '-') return usage(("Unrecognized argument: " + t).c_str());
        else { a.command = t; ++i; break; }
    }
    if (a.command == "depth") {
        for (; i < argc; ++i) {
            const std::string t = argv[i];
            if (t == "-d" || t == "--graph-depth-table") a.seg_depth = true;
            else if (t == "-r") { std::string v; if (!take(i, argc, argv, v)) return usage("No value provided for option '-r'."); a.paths.push_back(v); }
            else if (t == "-b" || t == "--bed-input") { if (!take(i, argc, argv, a.bed)) return usage("No value provided for option '-b'."); }
            else return usage(("Unrecognized argument: " + t).c_str());
        }
    } else if (a.command == "window-depth") {
        for (; i < argc; ++i) a.positional.push_back(argv[i]);
        if (a.positional.size() != 2) return usage("window-depth takes two positional arguments: <path> <window>");
    } else if (!a.command.empty()) {
        std::fprintf(stderr, "fgfa: subcommand '%s' is outside the scope of this build (node depth only)\n",
                     a.command.c_str());
        return 1;
    }
    // -m (cli/main.rs:59-65, 216-248): with only an output file it is the "preallocated" translation --
    // the GFA is parsed into a file whose pools have spare capacity (estimated from the text, or guessed
    // from -p when reading stdin).  With -i the file is opened for mutation in the reference; no command
    // of this build mutates, so it is simply viewed.
    if (a.mutate && a.command.empty() && a.input.empty() && !a.output.empty()) {
        try {
            flatgfa::HeapGFAStore store;
            flatgfa::file::Toc caps;
            if (!a.input_gfa.empty()) {
                flatgfa::MappedFile text(a.input_gfa);
                caps = flatgfa::estimate_toc(text.data(), text.size());
                store = flatgfa::Parser::parse_mem(text.data(), text.size());
            } else {
                caps = flatgfa::file::Toc::guess(a.prealloc_factor);
                store = flatgfa::Parser::parse_stream(stdin);
            }
            const std::vector<uint8_t> img = flatgfa::file::dump_preallocated(store.view(), caps);
            flatgfa::write_file(a.output, img.data(), img.size());
            return 0;
        } catch (const std::exception& e) {
            std::fprintf(stderr, "Error: %s\n", e.what());
            return 1;
        }
    }

    try {
        // Load the input from a file (binary) or stdin (text): main.rs:87-117.
        std::unique_ptr<flatgfa::MappedFile> map;
        flatgfa::HeapGFAStore store;
        flatgfa::FlatGFA gfa;
        if (!a.input.empty()) {
            map.reset(new flatgfa::MappedFile(a.input));                      // main.rs:99
            gfa = flatgfa::file::view_or_throw(map->data(), map->size());     // main.rs:100
        } else if (!a.input_gfa.empty()) {
            flatgfa::MappedFile text(a.input_gfa);                            // main.rs:106
            store = flatgfa::Parser::parse_mem(text.data(), text.size());     // main.rs:107
            gfa = store.view();
        } else {
            store = flatgfa::Parser::parse_stream(stdin);                     // main.rs:110-111
            gfa = store.view();
        }

        if (a.command == "depth") {                                           // main.rs:136-138
            if (!a.seg_depth && !a.bed.empty()) {                             // cmds.rs:246-255: interval depth table
                flatgfa::MappedFile file(a.bed);
                const flatgfa::HeapBEDStore bed = flatgfa::BEDParser::parse_mem(file.data(), file.size());
                auto depths = flatgfa::ops::window_depth::bed_depth(gfa, bed.view());
                flatgfa::ops::window_depth::IntervalDepth{bed.view(), std::move(depths)}.print();
                return 0;
            }
            if (!a.seg_depth) {                                               // cmds.rs:256-283: path depth table
                std::vector<uint32_t> ids;
                if (a.paths.empty()) {
                    for (size_t p = 0; p < gfa.paths.len(); ++p) ids.push_back((uint32_t)p);   // gfa.paths.ids()
                } else {
                    for (const auto& name : a.paths) {                        // find_path (flatgfa.rs:376-378); unknown names are dropped
                        for (size_t p = 0; p < gfa.paths.len(); ++p) {
                            const auto nm = gfa.get_path_name(gfa.paths[p]);
                            if (nm.len() == name.size() && std::memcmp(nm.data, name.data(), name.size()) == 0) {
                                ids.push_back((uint32_t)p);
                                break;
                            }
                        }
                    }
                }
                auto ld = flatgfa::ops::depth::path_depth(gfa, ids);
                flatgfa::ops::depth::PathDepth{gfa, std::move(ld.first), std::move(ld.second), ids}.print();
                return 0;
            }
            auto du = flatgfa::ops::depth::seg_depth_with_uniq(gfa, a.gpus);  // cmds.rs:239
            flatgfa::ops::depth::SegDepth{gfa, std::move(du.first), std::move(du.second)}.print();  // cmds.rs:240-245
            return 0;
        }

        if (a.command == "window-depth") {                                    // main.rs:178-180, cmds.rs:488-496
            const std::string& name = a.positional[0];
            char* endp = nullptr;
            const unsigned long long window = std::strtoull(a.positional[1].c_str(), &endp, 10);
            if (a.positional[1].empty() || *endp) return usage("window-depth: <window> must be a number");
            const int64_t path = gfa.find_path(reinterpret_cast<const uint8_t*>(name.data()), name.size());
            if (path < 0) throw flatgfa::Error("path not found");
            auto wd = flatgfa::ops::window_depth::window_depth(gfa, (uint32_t)path, window);
            flatgfa::ops::window_depth::IntervalDepth{wd.first.view(), std::move(wd.second)}.print();
            return 0;
        }

        // No command: emit the graph (main.rs:181-184, 190-213).
        if (!a.output.empty()) {
            std::vector<uint8_t> buf(flatgfa::file::size(gfa));
            flatgfa::file::dump(gfa, buf.data());
            flatgfa::write_file(a.output, buf.data(), buf.size());
            return 0;
        }
        std::string text;
        flatgfa::print::gfa(gfa, text);                                       // main.rs:201-211
        if (!a.output_gfa.empty()) {
            flatgfa::write_file(a.output_gfa, reinterpret_cast<const uint8_t*>(text.data()), text.size());
        } else {
            std::fwrite(text.data(), 1, text.size(), stdout);
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
}
