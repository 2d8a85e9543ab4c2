// filters.hpp -- SURVEY.md section 8(f4): the receive loop's message filters and the plain / CSV encoders, mirrored from
// rtlamr so that ertgpu_decode_file prints what `rtlamr -filterid=.. -filtertype=.. -unique=.. -format=plain|csv` would.
//   MessageFilter / FilterChain   protocol/parse.go:126-155
//   UintMap (flag syntax)          flags.go:200-224
//   MeterIDFilter, MeterTypeFilter flags.go:226-240
//   UniqueFilter                   flags.go:242-259 (meter id -> last checksum; applied BEFORE the digest of main.go:251-253)
//   LogMessage, PlainEncoder       protocol/parse.go:103-125, flags.go:261-272
//   csv.Encoder                    csv/csv.go:27-38 over Go's encoding/csv writer rules
// Header-only; nothing here touches a sample or the GPU.
#pragma once

#include <cstdint>
#include <cstdio>
#include <ctime>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "protocol.hpp"

namespace receiver {

struct MessageFilter {   // parse.go:153-155
    virtual ~MessageFilter() = default;
    virtual bool Filter(const protocol::Message& msg) = 0;
};

// flags.go:200-224: "1,2,3" -> set; anything that is not an unsigned decimal number is an error (strconv.ParseUint)
struct UintMap : std::set<uint64_t> {
    void Set(const std::string& value) {
        size_t i = 0;
        while (i <= value.size()) {
            size_t j = value.find(',', i);
            if (j == std::string::npos) j = value.size();
            const std::string v = value.substr(i, j - i);
            if (v.empty() || v.find_first_not_of("0123456789") != std::string::npos || v.size() > 20)
                throw std::invalid_argument("strconv.ParseUint: parsing \"" + v + "\": invalid syntax");
            insert(std::stoull(v));
            i = j + 1;
        }
    }
};

struct MeterIDFilter : MessageFilter {   // flags.go:226-232
    UintMap ids;
    bool Filter(const protocol::Message& msg) override { return ids.count(msg.MeterID()) != 0; }
};

struct MeterTypeFilter : MessageFilter {   // flags.go:234-240
    UintMap types;
    bool Filter(const protocol::Message& msg) override { return types.count(msg.MeterType()) != 0; }
};

struct UniqueFilter : MessageFilter {   // flags.go:242-259
    std::map<uint32_t, std::vector<uint8_t>> last;
    bool Filter(const protocol::Message& msg) override {
        const std::vector<uint8_t> checksum = msg.Checksum();
        const uint32_t mid = msg.MeterID();
        auto it = last.find(mid);
        if (it != last.end() && it->second == checksum) return false;
        last[mid] = checksum;
        return true;
    }
};

struct FilterChain {   // parse.go:126-149
    std::vector<std::shared_ptr<MessageFilter>> filters;
    void Add(std::shared_ptr<MessageFilter> f) { filters.push_back(std::move(f)); }
    bool Match(const protocol::Message& msg) {
        for (auto& f : filters)
            if (!f->Filter(msg)) return false;   // later filters are not consulted (a rejected message never reaches -unique)
        return true;
    }
    bool empty() const { return filters.empty(); }
};

// parse.go:103-125.  Time is kept as (seconds, nanoseconds) UTC; the reference formats local time with the same layouts.
struct LogMessage {
    int64_t unix_seconds = 0;
    int32_t nanos = 0;
    int64_t Offset = 0;
    int Length = 0;
    const protocol::Message* Message = nullptr;

    std::string time_layout(bool rfc3339nano) const {
        std::tm tm{};
        const time_t t = (time_t)unix_seconds;
        gmtime_r(&t, &tm);
        char buf[64];
        if (!rfc3339nano) {   // TimeFormat = "2006-01-02T15:04:05.000" (parse.go:13)
            snprintf(buf, sizeof(buf), "%04d-%02d-%02dT%02d:%02d:%02d.%03d", tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday, tm.tm_hour,
                     tm.tm_min, tm.tm_sec, nanos / 1000000);
            return buf;
        }
        // time.RFC3339Nano: fractional seconds with trailing zeros removed, "Z" for UTC
        snprintf(buf, sizeof(buf), "%04d-%02d-%02dT%02d:%02d:%02d", tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday, tm.tm_hour,
                 tm.tm_min, tm.tm_sec);
        std::string s = buf;
        if (nanos) {
            char f[16];
            snprintf(f, sizeof(f), ".%09d", nanos);
            std::string fs = f;
            while (fs.back() == '0') fs.pop_back();
            s += fs;
        }
        return s + "Z";
    }
    std::string String() const {   // parse.go:113-117
        return "{Time:" + time_layout(false) + " Offset:" + std::to_string(Offset) + " Length:" + std::to_string(Length) + " " +
               Message->MsgType() + ":" + Message->String() + "}";
    }
    std::string StringNoOffset() const {   // parse.go:119-121
        return "{Time:" + time_layout(false) + " " + Message->MsgType() + ":" + Message->String() + "}";
    }
    std::vector<std::string> Record() const {   // parse.go:123-129
        std::vector<std::string> r = {time_layout(true), std::to_string(Offset), std::to_string(Length)};
        for (auto& f : Message->Record()) r.push_back(f);
        return r;
    }
};

// flags.go:261-272: StringNoOffset when the sample file is os.DevNull (the default), String otherwise
inline std::string PlainLine(const LogMessage& m, bool sample_file_is_devnull) {
    return sample_file_is_devnull ? m.StringNoOffset() : m.String();
}

// One record in Go's encoding/csv writer format (Comma = ',', UseCRLF = false), csv/csv.go:36: a field is quoted when it
// is `\.`, contains the delimiter, a quote, CR or LF, or starts with a space; quotes are doubled; "" stays unquoted.
inline std::string CsvLine(const std::vector<std::string>& fields) {
    std::string out;
    for (size_t n = 0; n < fields.size(); n++) {
        const std::string& f = fields[n];
        if (n) out += ',';
        bool quote = false;
        if (!f.empty()) {
            quote = f == "\\." || f.find_first_of(",\"\r\n") != std::string::npos || f[0] == ' ' || f[0] == '\t';
        }
        if (!quote) {
            out += f;
            continue;
        }
        out += '"';
        for (char c : f) {
            if (c == '"') out += "\"\"";
            else out += c;   // UseCRLF = false: CR and LF are written as they are
        }
        out += '"';
    }
    return out + "\n";
}

}  // namespace receiver
