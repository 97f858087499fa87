// GFA text output of a FlatGFA: C++ restatement of the reference's flatgfa/src/print.rs
// (`Display` for Handle :39-45, Path :47-66, Link :68-87, Segment :89-98, the whole graph
// :100-152).  Needed by the format round trip (`fgfa < x.gfa`, `fgfa -i x.flatgfa`,
// tests/turnt.toml:162-172) that pins the parser and the .flatgfa writer.
#pragma once
#include <string>

#include "flatgfa.hpp"

namespace flatgfa {
namespace print {

void handle(const FlatGFA& gfa, Handle h, std::string& out);          // "12+"
void segment(const FlatGFA& gfa, const Segment& s, std::string& out); // "S\t12\tACGT[\toptional]"
void path(const FlatGFA& gfa, const Path& p, std::string& out);       // "P\tname\t1+,2-\t*"
void link(const FlatGFA& gfa, const Link& l, std::string& out);       // "L\t1\t+\t2\t-\t0M"

// print.rs:144-152: the original line order if one was recorded, else header, segments,
// paths, links.  Every line ends with '\n'.
void gfa(const FlatGFA& gfa, std::string& out);

}  // namespace print
}  // namespace flatgfa
