#include "print.hpp"

namespace flatgfa {
namespace print {

namespace {
void bytes(const Pool<uint8_t>& p, std::string& out) { out.append(reinterpret_cast<const char*>(p.data), p.len()); }

// print.rs:13-34: opcode letters follow the enum order Match, Gap, Insertion, Deletion -> M N D I;
// an empty alignment prints as "0M".
void alignment(const FlatGFA& g, Span ops, std::string& out) {
    const Pool<AlignOp> a = g.alignment.slice(ops);
    if (a.is_empty()) out += "0M";
    for (const AlignOp& op : a) {
        out += std::to_string(op.bits >> 8);
        const uint32_t code = op.bits & 0xFFu;
        if (code > 3) throw Error("invalid alignment opcode");
        out += "MNDI"[code];
    }
}
}  // namespace

void handle(const FlatGFA& g, Handle h, std::string& out) {            // print.rs:39-45
    out += std::to_string((uint64_t)g.segs[h.segment()].name);
    out += h.is_forward() ? '+' : '-';
}

void segment(const FlatGFA& g, const Segment& s, std::string& out) {   // print.rs:89-98
    out += "S\t";
    out += std::to_string((uint64_t)s.name);
    out += '\t';
    bytes(g.get_seq(s), out);
    if (!s.optional.is_empty()) {
        out += '\t';
        bytes(g.optional_data.slice(s.optional), out);
    }
}

void path(const FlatGFA& g, const Path& p, std::string& out) {         // print.rs:47-66
    out += "P\t";
    bytes(g.get_path_name(p), out);
    out += '\t';
    const Pool<Handle> steps = g.get_path_steps(p);
    if (steps.is_empty()) throw Error("index out of bounds: the len is 0 but the index is 0");   // `steps[0]`, :51
    for (size_t i = 0; i < steps.len(); ++i) {
        if (i) out += ',';
        handle(g, steps.data[i], out);
    }
    out += '\t';
    const Pool<Span> ov = g.overlaps.slice(p.overlaps);
    if (ov.is_empty()) {
        out += '*';
    } else {
        for (size_t i = 0; i < ov.len(); ++i) {
            if (i) out += ',';
            alignment(g, ov.data[i], out);
        }
    }
}

void link(const FlatGFA& g, const Link& l, std::string& out) {         // print.rs:68-87
    out += "L\t";
    out += std::to_string((uint64_t)g.segs[l.from.segment()].name);
    out += l.from.is_forward() ? "\t+\t" : "\t-\t";
    out += std::to_string((uint64_t)g.segs[l.to.segment()].name);
    out += l.to.is_forward() ? "\t+\t" : "\t-\t";
    alignment(g, l.overlap, out);
}

void gfa(const FlatGFA& g, std::string& out) {
    auto header = [&] {
        out += "H\t";
        bytes(g.header, out);
        out += '\n';
    };
    if (g.line_order.is_empty()) {                                      // write_normalized, print.rs:129-142
        if (!g.header.is_empty()) header();
        for (const Segment& s : g.segs) { segment(g, s, out); out += '\n'; }
        for (const Path& p : g.paths) { path(g, p, out); out += '\n'; }
        for (const Link& l : g.links) { link(g, l, out); out += '\n'; }
        return;
    }
    size_t si = 0, pi = 0, li = 0;                                      // write_preserved, print.rs:100-127
    for (const uint8_t kind : g.line_order) {
        switch (kind) {
            case kLineHeader:
                if (g.header.is_empty()) throw Error("assertion failed: !version.is_empty()");
                header();
                break;
            case kLineSegment:
                if (si >= g.segs.len()) throw Error("too few segments");
                segment(g, g.segs.data[si++], out);
                out += '\n';
                break;
            case kLinePath:
                if (pi >= g.paths.len()) throw Error("too few paths");
                path(g, g.paths.data[pi++], out);
                out += '\n';
                break;
            case kLineLink:
                if (li >= g.links.len()) throw Error("too few links");
                link(g, g.links.data[li++], out);
                out += '\n';
                break;
            default:
                throw Error("invalid line kind");                       // `try_into().unwrap()`, flatgfa.rs:418-423
        }
    }
}

}  // namespace print
}  // namespace flatgfa
