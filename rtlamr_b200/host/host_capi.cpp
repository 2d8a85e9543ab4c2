// host_capi.cpp -- a small C surface over the C++ protocol mirror so that the Python parity tests
// (and a CLI) can drive Decoder/Parser exactly like rtlamr's main.go does (main.go:59-86,235).
#include <cstdio>
#include <cstring>
#include <memory>
#include <sstream>

#include "protocol.hpp"
#include "receiver.hpp"

extern "C" {

struct erthost_msg {
    int64_t block;
    int32_t idx;
    uint32_t meter_id;
    uint32_t meter_type;
    int32_t nchecksum;
    uint8_t checksum[8];
    char msgtype[12];
    char text[1400];    // Message.String()
    char record[1400];  // Message.Record() joined with ','
};

struct erthost {
    protocol::Decoder d;
    std::string err;
};

// main.go:59-86: NewDecoder; for each -msgtype NewParser + RegisterProtocol; Allocate
erthost* erthost_new(const char* msgtypes_csv, int chip_length, int device, long long max_blocks, long long max_cands,
                     char* errbuf, int errcap) {
    std::unique_ptr<erthost> h;   // released only on success: a throwing NewParser / Allocate must not leak the handle
    try {
        protocol::RegisterStockParsers();
        h.reset(new erthost());
        std::stringstream ss(msgtypes_csv);
        std::string name;
        while (std::getline(ss, name, ',')) {
            if (name.empty()) continue;
            h->d.RegisterProtocol(protocol::NewParser(name, chip_length));
        }
        h->d.Allocate(device, max_blocks, max_cands);
        return h.release();
    } catch (const std::exception& e) {
        if (errbuf && errcap > 0) snprintf(errbuf, (size_t)errcap, "%s", e.what());
        return nullptr;
    }
}

// The same without Allocate: no device is touched, only erthost_parse works on such a handle.
erthost* erthost_new_parse_only(const char* msgtypes_csv, int chip_length, char* errbuf, int errcap) {
    std::unique_ptr<erthost> h;
    try {
        protocol::RegisterStockParsers();
        h.reset(new erthost());
        std::stringstream ss(msgtypes_csv);
        std::string name;
        while (std::getline(ss, name, ',')) {
            if (name.empty()) continue;
            h->d.RegisterProtocol(protocol::NewParser(name, chip_length));
        }
        return h.release();
    } catch (const std::exception& e) {
        if (errbuf && errcap > 0) snprintf(errbuf, (size_t)errcap, "%s", e.what());
        return nullptr;
    }
}

void erthost_free(erthost* h) { delete h; }

const char* erthost_error(const erthost* h) { return h ? h->err.c_str() : ""; }

int erthost_config(const erthost* h, int32_t* out12) {
    const protocol::PacketConfig& c = h->d.Cfg;
    const int32_t v[12] = {c.DataRate, c.BlockSize, c.BlockSize2, c.ChipLength, c.SymbolLength, c.SampleRate,
                           c.PreambleSymbols, c.PacketSymbols, c.PreambleLength, c.PacketLength, c.BufferLength,
                           (int32_t)c.CenterFreq};
    memcpy(out12, v, sizeof(v));
    return 0;
}

int erthost_reset(erthost* h) {
    try {
        h->d.Reset();
        return 0;
    } catch (const std::exception& e) {
        h->err = e.what();
        return -1;
    }
}

// rcvr.d.Decode(block) for N blocks (main.go:235); returns the number of messages (all are counted,
// the first `cap` are written), or -1 on error / -2 when len is not a multiple of BlockSize2.
static long long fill_msgs(const std::vector<protocol::MessagePtr>& msgs, erthost_msg* out, long long cap);

// The parsers alone (decode.go:177-187 + each Parse): candidates sorted by (block, preamble, idx) in, messages out.
long long erthost_parse(erthost* h, const ertgpu_candidate* cands, long long n, erthost_msg* out, long long cap) {
    try {
        std::vector<protocol::MessagePtr> msgs;
        h->d.Dispatch(cands, (size_t)n, msgs);
        return fill_msgs(msgs, out, cap);
    } catch (const std::exception& e) {
        h->err = e.what();
        return -1;
    }
}

// erthost_parse followed by the receive loop's cross-block dedup (main.go:244-260,292) with a fresh memory:
// the messages rtlamr would print for this candidate list with -blockdedup=`unique`.
long long erthost_parse_dedup(erthost* h, const ertgpu_candidate* cands, long long n, int block_dedup, erthost_msg* out,
                              long long cap, long long* duplicates) {
    try {
        std::vector<protocol::MessagePtr> msgs, kept;
        h->d.Dispatch(cands, (size_t)n, msgs);
        receiver::BlockDedup dd;
        receiver::Stats st;
        std::vector<const protocol::Message*> order;
        dd.Filter(msgs, block_dedup != 0, [&](const protocol::Message& m) { order.push_back(&m); }, st);
        if (duplicates) *duplicates = st.duplicates;
        // re-own in emit order (msgs keeps the objects alive until we return)
        std::vector<protocol::MessagePtr> view;
        for (const protocol::Message* m : order)
            for (auto& p : msgs)
                if (p.get() == m) {
                    view.push_back(std::move(p));
                    break;
                }
        return fill_msgs(view, out, cap);
    } catch (const std::exception& e) {
        h->err = e.what();
        return -1;
    }
}

// erthost_parse, then the receive loop of main.go:236-260 on the result with a fresh memory: the filter chain in the order
// flag.Visit builds it (main.go:97-113: filterid, filtertype, unique -- lexicographic, whatever the command line's order),
// then the cross-block digest dedup.  filterid / filtertype: comma lists as for the flags, NULL or "" = flag not given.
long long erthost_parse_filtered(erthost* h, const ertgpu_candidate* cands, long long n, const char* filterid, const char* filtertype,
                                 int unique, int block_dedup, erthost_msg* out, long long cap, long long* duplicates,
                                 long long* filtered) {
    try {
        std::vector<protocol::MessagePtr> msgs;
        h->d.Dispatch(cands, (size_t)n, msgs);
        receiver::FilterChain fc;
        if (filterid && *filterid) {
            auto f = std::make_shared<receiver::MeterIDFilter>();
            f->ids.Set(filterid);
            fc.Add(f);
        }
        if (filtertype && *filtertype) {
            auto f = std::make_shared<receiver::MeterTypeFilter>();
            f->types.Set(filtertype);
            fc.Add(f);
        }
        if (unique) fc.Add(std::make_shared<receiver::UniqueFilter>());
        receiver::BlockDedup dd;
        receiver::Stats st;
        std::vector<const protocol::Message*> order;
        dd.Filter(msgs, block_dedup != 0, [&](const protocol::Message& m) { order.push_back(&m); }, st, fc.empty() ? nullptr : &fc);
        if (duplicates) *duplicates = st.duplicates;
        if (filtered) *filtered = st.filtered;
        std::vector<protocol::MessagePtr> view;
        for (const protocol::Message* m : order)
            for (auto& p : msgs)
                if (p.get() == m) {
                    view.push_back(std::move(p));
                    break;
                }
        return fill_msgs(view, out, cap);
    } catch (const std::exception& e) {
        h->err = e.what();
        return -1;
    }
}

// The encoders (flags.go:140-151) on every message of erthost_parse: format 0 = plain (PlainEncoder, flags.go:261-272),
// 1 = csv (csv/csv.go:27-38); the LogMessage fields (parse.go:103-109) are the caller's.  Returns the number of bytes the
// lines need (they are written when they fit `cap`), or -1.
long long erthost_encode(erthost* h, const ertgpu_candidate* cands, long long n, int format, long long unix_seconds, int nanos,
                         long long offset, int length, int sample_file_is_devnull, char* buf, long long cap) {
    try {
        std::vector<protocol::MessagePtr> msgs;
        h->d.Dispatch(cands, (size_t)n, msgs);
        std::string all;
        for (auto& m : msgs) {
            receiver::LogMessage lm;
            lm.unix_seconds = unix_seconds;
            lm.nanos = nanos;
            lm.Offset = offset;
            lm.Length = length;
            lm.Message = m.get();
            all += format == 1 ? receiver::CsvLine(lm.Record()) : receiver::PlainLine(lm, sample_file_is_devnull != 0) + "\n";
        }
        if ((long long)all.size() < cap && buf) memcpy(buf, all.c_str(), all.size() + 1);
        return (long long)all.size();
    } catch (const std::exception& e) {
        h->err = e.what();
        return -1;
    }
}

long long erthost_decode(erthost* h, const uint8_t* iq, size_t len, erthost_msg* out, long long cap) {
    try {
        auto msgs = h->d.Decode(iq, len);
        return fill_msgs(msgs, out, cap);
    } catch (const std::length_error& e) {
        h->err = e.what();
        return -2;
    } catch (const std::exception& e) {
        h->err = e.what();
        return -1;
    }
}

static long long fill_msgs(const std::vector<protocol::MessagePtr>& msgs, erthost_msg* out, long long cap) {
    {
        long long n = 0;
        for (auto& m : msgs) {
            if (n < cap) {
                erthost_msg& o = out[n];
                memset(&o, 0, sizeof(o));
                o.block = m->Block;
                o.idx = m->Idx;
                o.meter_id = m->MeterID();
                o.meter_type = m->MeterType();
                auto ck = m->Checksum();
                o.nchecksum = (int32_t)ck.size();
                memcpy(o.checksum, ck.data(), std::min<size_t>(ck.size(), 8));
                snprintf(o.msgtype, sizeof(o.msgtype), "%s", m->MsgType().c_str());
                snprintf(o.text, sizeof(o.text), "%s", m->String().c_str());
                std::string rec;
                for (auto& f : m->Record()) rec += (rec.empty() ? "" : ",") + f;
                snprintf(o.record, sizeof(o.record), "%s", rec.c_str());
            }
            n++;
        }
        return n;
    }
}

int erthost_log(const erthost* h, char* buf, int cap) {
    snprintf(buf, (size_t)cap, "%s", h->d.Log().c_str());
    return 0;
}

}  // extern "C"
