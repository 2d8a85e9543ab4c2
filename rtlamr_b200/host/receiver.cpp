// receiver.cpp -- see receiver.hpp.
#include "receiver.hpp"

#include <chrono>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace receiver {

Digest NewDigest(const protocol::Message& m) { return Digest(m.MsgType(), m.MeterType(), m.MeterID(), m.Checksum()); }

Receiver::Receiver(const std::string& msgtypes, int chipLength, int device, int64_t blocksPerCall) {
    protocol::RegisterStockParsers();
    std::string list = msgtypes == "all" ? "scm,scm+,idm,r900" : msgtypes;  // main.go:67-73
    std::stringstream ss(list);
    std::string name;
    while (std::getline(ss, name, ','))
        if (!name.empty()) d_.RegisterProtocol(protocol::NewParser(name, chipLength));  // main.go:77-82
    if (blocksPerCall <= 0) blocksPerCall = 4096;
    blocks_per_call_ = blocksPerCall;
    d_.Allocate(device, blocksPerCall, 0);  // main.go:86
    buf_bytes_ = (size_t)blocksPerCall * (size_t)d_.Cfg.BlockSize2;
    for (int k = 0; k < 2; k++) {
        void* p = nullptr;
        if (ertgpu_host_alloc(&p, buf_bytes_) != ERTGPU_OK) {
            for (int j = 0; j < k; j++) {   // a constructor that throws runs no destructor: give back what it got
                ertgpu_host_free(buf_[j]);
                buf_[j] = nullptr;
            }
            throw std::runtime_error("pinned host allocation failed");
        }
        buf_[k] = static_cast<uint8_t*>(p);
    }
}

Receiver::~Receiver() {
    for (int k = 0; k < 2; k++) ertgpu_host_free(buf_[k]);
}

void Receiver::Filter(std::vector<protocol::MessagePtr>& msgs, bool block_dedup,
                      const std::function<void(const protocol::Message&)>& emit, Stats& st) {
    dedup_.Filter(msgs, block_dedup, emit, st, fc_.empty() ? nullptr : &fc_);
}

void BlockDedup::Filter(std::vector<protocol::MessagePtr>& msgs, bool block_dedup,
                        const std::function<void(const protocol::Message&)>& emit, Stats& st, FilterChain* fc) {
    // main.go:221-224,244-260,292: `next` collects the digests of the current block, a message whose digest
    // was seen in the previous block is skipped, and the maps swap after every block -- including blocks
    // without messages, which is why a gap in the block numbers empties `prev`.
    size_t i = 0;
    while (i < msgs.size()) {
        const int64_t b = msgs[i]->Block;
        if (b != prev_block_ + 1) prev_.clear();
        std::set<Digest> next;
        for (; i < msgs.size() && msgs[i]->Block == b; i++) {
            if (fc && !fc->Match(*msgs[i])) {   // main.go:236-239: rejected before the digest is recorded
                st.filtered++;
                continue;
            }
            Digest dg = NewDigest(*msgs[i]);
            next.insert(dg);
            if (block_dedup && prev_.count(dg)) {
                st.duplicates++;
                continue;
            }
            st.messages++;
            emit(*msgs[i]);
        }
        prev_.swap(next);
        prev_block_ = b;
    }
}

void Receiver::Reset() {
    d_.Reset();
    dedup_.Reset();
}

Stats Receiver::Run(FILE* in, bool block_dedup, const std::function<void(const protocol::Message&)>& emit) {
    Stats st;
    const size_t bs2 = (size_t)d_.Cfg.BlockSize2;
    const auto t0 = std::chrono::steady_clock::now();
    int k = 0;
    size_t carry = 0;  // bytes of a partial block kept at the front of the next buffer
    for (;;) {
        size_t have = carry;
        while (have < buf_bytes_) {
            const size_t n = fread(buf_[k] + have, 1, buf_bytes_ - have, in);
            if (n == 0) break;
            have += n;
        }
        const size_t whole = have / bs2 * bs2;
        if (whole) {
            auto msgs = d_.Decode(buf_[k], whole);
            st.blocks += (int64_t)(whole / bs2);
            st.bytes += (int64_t)whole;
            Filter(msgs, block_dedup, emit, st);
        }
        carry = have - whole;
        if (carry) memcpy(buf_[k ^ 1], buf_[k] + whole, carry);
        k ^= 1;
        if (have < buf_bytes_) break;  // EOF: the trailing partial block is dropped like main.go does
    }
    st.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return st;
}

}  // namespace receiver
