// fkb_loader.h -- host sequence loader: record/line stripping per the stream contract (internal).
// Restates what findKmer()'s outer loop does with '\n', '>' and byte 0xFF
// (reference findKmer/src/findKmer.cpp:988-1011) as a block-parallel filter.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace fkb {

struct StripResult {
    size_t n_out;        // bytes written
    size_t stop_pos;     // position of the first 0xFF met outside a header inside the block, or SIZE_MAX
    bool ends_in_header; // the block ends inside an unterminated '>' line
};

// True iff position `pos` of buf lies inside a '>' header line (a '>' occurs between the previous '\n' and pos).
bool in_header_at(const uint8_t *buf, size_t pos);

// in_header_at() for every block start first + i * block_bytes of buf[first, len), in one bounded pass (see fkb_loader.cpp)
std::vector<uint8_t> header_states(const uint8_t *buf, size_t first, size_t len, size_t block_bytes, int n_threads);

// Strip buf[a,b) into out (room for b-a bytes).  `in_header`: state at a (from in_header_at).
// Stops at the first 0xFF outside a header.
StripResult strip_block(const uint8_t *buf, size_t a, size_t b, bool in_header, uint8_t *out);

// Same scan without writing: only counts (first pass of the two-pass parallel strip).
StripResult strip_block_count(const uint8_t *buf, size_t a, size_t b, bool in_header);

int default_host_threads();

}  // namespace fkb
