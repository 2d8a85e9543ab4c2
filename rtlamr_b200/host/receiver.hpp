// receiver.hpp -- SURVEY.md section 8(f1): file/stream ingest in place of rtl_tcp, plus the receive loop's
// cross-block message dedup.  Mirrors the parts of rtlamr's main.go Receiver that sit around
// protocol.Decoder.Decode:
//   * the block reader (main.go:156-205) -- here it reads K*BlockSize2 bytes per iteration from a file or
//     stdin into pinned host buffers (ertgpu_host_alloc) instead of one block from a TCP socket;
//   * the decode loop (main.go:207-294) -- Decode, then the prev/next digest maps that drop a message
//     already reported for the previous block (main.go:244-260,292).
// Whole blocks only: a trailing partial block is never decoded (main.go:166-186).
#pragma once

#include <cstdio>
#include <functional>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "filters.hpp"
#include "protocol.hpp"

namespace receiver {

// protocol.Digest (parse.go:87-101)
using Digest = std::tuple<std::string, uint8_t, uint32_t, std::vector<uint8_t>>;
Digest NewDigest(const protocol::Message& m);

struct Stats {
    int64_t blocks = 0, bytes = 0, messages = 0, duplicates = 0, filtered = 0;
    double seconds = 0;
};

// The receive loop's prev/next digest maps (main.go:221-224,244-260,292) on their own: a message whose digest was
// reported for the previous block is dropped; a gap in the block numbers empties the memory.
class BlockDedup {
public:
    // msgs grouped by ascending Block (what Decode returns); `emit` sees every message that is reported.  `fc` is the
    // receive loop's filter chain (main.go:236-239): a message it rejects is skipped BEFORE its digest is recorded.
    void Filter(std::vector<protocol::MessagePtr>& msgs, bool block_dedup,
                const std::function<void(const protocol::Message&)>& emit, Stats& st, FilterChain* fc = nullptr);
    void Reset() {
        prev_.clear();
        prev_block_ = -2;
    }

private:
    std::set<Digest> prev_;
    int64_t prev_block_ = -2;
};

class Receiver {
public:
    // msgtypes: comma list as for -msgtype ("all" = scm,scm+,idm,r900, main.go:67-73)
    Receiver(const std::string& msgtypes, int chipLength, int device, int64_t blocksPerCall);
    ~Receiver();

    protocol::Decoder& decoder() { return d_; }
    FilterChain& filters() { return fc_; }   // -filterid / -filtertype / -unique (flags.go:147-170)

    // Run over a FILE* until EOF.  `emit` is called once per reported message (after the cross-block dedup when block_dedup).
    Stats Run(FILE* in, bool block_dedup, const std::function<void(const protocol::Message&)>& emit);

    // Start a new, independent stream (the next file of a batch): zeroed Decoder history (decode.go:144-145),
    // block numbers restart at 0, empty dedup set.
    void Reset();

    // The dedup step on one Decode result (messages grouped by ascending Block), exposed for tests.
    void Filter(std::vector<protocol::MessagePtr>& msgs, bool block_dedup,
                const std::function<void(const protocol::Message&)>& emit, Stats& st);

private:
    protocol::Decoder d_;
    int64_t blocks_per_call_;
    uint8_t* buf_[2] = {nullptr, nullptr};
    size_t buf_bytes_ = 0;
    BlockDedup dedup_;
    FilterChain fc_;
};

}  // namespace receiver
