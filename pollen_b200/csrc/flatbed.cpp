#include "flatbed.hpp"

#include <cstring>

namespace flatgfa {

namespace {
// atoi 2.0.0 `FromRadix10::from_radix_10` for u64 (flatbed.rs:110-116 parse_num): leading ASCII
// digits only, no sign, wrapping on overflow (release build); `used == 0` is the error.
size_t parse_u64(const uint8_t* s, size_t n, uint64_t* out) {
    uint64_t v = 0;
    size_t i = 0;
    while (i < n && s[i] >= '0' && s[i] <= '9') { v = v * 10 + (uint64_t)(s[i] - '0'); ++i; }
    *out = v;
    return i;
}

void parse_line(HeapBEDStore& flat, const uint8_t* line, size_t n) {   // flatbed.rs:141-152
    if (n && line[0] == '#') return;                                    // :143-145
    const uint8_t* tab = static_cast<const uint8_t*>(std::memchr(line, '\t', n));   // gfaline.rs:129-142 parse_field
    const size_t name_len = tab ? (size_t)(tab - line) : n;
    const uint8_t* rest = tab ? tab + 1 : line + n;
    size_t rest_len = tab ? n - name_len - 1 : 0;
    uint64_t start = 0, end = 0;
    size_t used = parse_u64(rest, rest_len, &start);                    // :148
    if (used == 0) throw Error("expected number");
    rest += used;
    rest_len -= used;
    if (rest_len == 0) throw Error("range start index 1 out of range for slice of length 0");   // `&rest[1..]`, :149
    used = parse_u64(rest + 1, rest_len - 1, &end);
    if (used == 0) throw Error("expected number");
    flat.add_entry(line, name_len, start, end);                         // :151
}
}  // namespace

HeapBEDStore BEDParser::parse_mem(const uint8_t* buf, size_t len) {     // flatbed.rs:126-131
    HeapBEDStore flat;
    size_t pos = 0;
    while (pos < len) {
        const uint8_t* nl = static_cast<const uint8_t*>(std::memchr(buf + pos, '\n', len - pos));
        if (!nl) break;                                                 // memfile.rs:59: no needle, no line
        parse_line(flat, buf + pos, (size_t)(nl - (buf + pos)));
        pos = (size_t)(nl - buf) + 1;
    }
    return flat;
}

}  // namespace flatgfa
