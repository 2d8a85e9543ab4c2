// protocol.hpp -- host side above the C ABI: a C++ mirror of rtlamr's `protocol` package surface
// for the Decoder hot path (reference protocol/decode.go, protocol/parse.go) and of the per-message
// parsers that consume its output (scm/, scmplus/, idm/, netidm/, r900/, r900bcd/).
//
// The reference is Go and no Go toolchain exists in this image, so this mirror stands in for the
// cgo shim (go/protocol/decode_cuda.go, shown in INTEGRATION.md) in the parity tests: same names,
// same argument meaning, same error behaviour.  All DSP runs in libertgpu.so (CUDA); the parsers
// below are the reference's pure byte-level logic (CRC re-check, field extraction) and never
// touch a sample.
//
// Differences forced by the language, not by design:
//   * Parse returns messages through a std::vector instead of a Go channel + WaitGroup
//     (parse.go:72-76); message order inside a block is unspecified in the reference
//     (map iteration + goroutines, decode.go:177-187), here it is registration order.
//   * Decode accepts N*BlockSize2 bytes and is equivalent to N sequential reference Decode calls.
#pragma once

#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ertgpu.h"

namespace crc {  // reference crc/crc.go

using Table = std::array<uint16_t, 256>;
Table NewTable(uint16_t poly);                                                   // crc.go:34-47
uint16_t Checksum(uint16_t init, const uint8_t* data, size_t n, const Table& t); // crc.go:49-55

struct CRC {  // crc.go:7-14
    std::string Name;
    uint16_t Init = 0, Poly = 0, Residue = 0;
    Table tbl{};
    uint16_t Checksum(const uint8_t* data, size_t n) const { return crc::Checksum(Init, data, n, tbl); }  // crc.go:28-30
    std::string String() const;                                                                         // crc.go:24-26
};
CRC NewCRC(const std::string& name, uint16_t init, uint16_t poly, uint16_t residue);  // crc.go:16-22

}  // namespace crc

namespace gf {  // reference r900/gf/gf.go, NewField(32, 37, 2) only

struct Field {
    std::array<uint8_t, 62> exp{};
    std::array<uint8_t, 32> log{};
    Field();                                                                            // gf.go:20-57
    uint8_t Mul(uint8_t x, uint8_t y) const;                                            // gf.go:143-148
    std::vector<uint8_t> Syndrome(const uint8_t* message, int n, int parity, int offset) const;  // gf.go:152-172
};

}  // namespace gf

namespace protocol {

// decode.go:27-42
struct PacketConfig {
    std::string Protocol, Preamble;
    int DataRate = 0;
    int BlockSize = 0, BlockSize2 = 0;
    int ChipLength = 0, SymbolLength = 0;
    int SampleRate = 0;
    int PreambleSymbols = 0, PacketSymbols = 0;
    int PreambleLength = 0, PacketLength = 0;
    int BufferLength = 0;
    uint32_t CenterFreq = 0;
};

// parse.go:55-59 plus what the GPU attaches to each candidate
struct Data {
    int Idx = 0;
    std::string Bits;
    std::vector<uint8_t> Bytes;
    int64_t Block = 0;                 // which reference Decode call produced it
    uint32_t CheckMask = 0;            // GPU integrity screens (ertgpu_candidate.check_mask)
    bool HasR900 = false;
    std::array<uint8_t, ERTGPU_R900_DIGITS> R900Digits{};  // r900 quantized[] at the payload positions
};
Data NewData(const uint8_t* data, size_t n);  // parse.go:61-69

// parse.go:78-84 (csv.Recorder folded in)
struct Message {
    virtual ~Message() = default;
    virtual std::string MsgType() const = 0;
    virtual uint32_t MeterID() const = 0;
    virtual uint8_t MeterType() const = 0;
    virtual std::vector<uint8_t> Checksum() const = 0;
    virtual std::string String() const = 0;
    virtual std::vector<std::string> Record() const = 0;
    int64_t Block = 0;  // bookkeeping for the parity tests (not in the reference)
    int Idx = 0;
};
using MessagePtr = std::unique_ptr<Message>;

class Decoder;

// parse.go:72-76
struct Parser {
    virtual ~Parser() = default;
    virtual void Parse(const std::vector<Data>& pkts, std::vector<MessagePtr>& msgCh) = 0;
    virtual void SetDecoder(Decoder*) {}
    virtual PacketConfig Cfg() const = 0;
    // the GPU-side screen equivalent to this parser's own integrity check (ertgpu_protocol.check_*)
    virtual void Screen(ertgpu_protocol& p) const { p.check_kind = ERTGPU_CHECK_NONE; }
};
using ParserPtr = std::shared_ptr<Parser>;
using NewParserFunc = std::function<ParserPtr(int chipLength)>;  // parse.go:21

void RegisterParser(const std::string& name, NewParserFunc fn);  // parse.go:28-39 (throws like Go panics)
ParserPtr NewParser(const std::string& name, int chipLength);    // parse.go:42-51 (throws invalid_argument)

// decode.go:45-63.  Signal/Quantized are not materialised on the host: the parity taps of the C ABI
// expose them (ertgpu_tap).
class Decoder {
public:
    PacketConfig Cfg;

    Decoder();   // NewDecoder, decode.go:65-71
    ~Decoder();
    Decoder(const Decoder&) = delete;
    Decoder& operator=(const Decoder&) = delete;

    void RegisterProtocol(ParserPtr p);  // decode.go:100-128
    void Allocate(int device = 0, int64_t maxBlocksPerCall = 0, int64_t maxCandidates = 0);  // decode.go:131-160
    // decode.go:163-197 for N = len/BlockSize2 consecutive blocks; throws std::length_error when len is
    // not a multiple of BlockSize2 (the reference panics with a slice-bounds error, decode.go:222)
    std::vector<MessagePtr> Decode(const uint8_t* input, size_t len);
    // The second half of Decode (decode.go:177-187 + parse.go): candidates sorted by (block, preamble, idx) are
    // regrouped into per-block, per-preamble []Data and handed to the parsers filed under that preamble.  Needs
    // RegisterProtocol only (no device), so the parsers can be exercised on candidates from any source.
    void Dispatch(const ertgpu_candidate* cands, size_t n, std::vector<MessagePtr>& out);
    std::string Log() const;  // decode.go:73-90
    void Reset();
    ertgpu_handle* Handle() const { return h_; }

private:
    ertgpu_handle* h_ = nullptr;
    std::vector<ParserPtr> parsers_;            // registration order
    std::vector<std::string> preambles_;        // distinct, registration order
    std::vector<std::vector<size_t>> by_pre_;   // parser indices per preamble (decode.go:124)
    std::vector<ertgpu_candidate> cands_;
    bool allocated_ = false;
};

void RegisterStockParsers();  // the effect of the reference's blank imports (main.go:37-44)

}  // namespace protocol
