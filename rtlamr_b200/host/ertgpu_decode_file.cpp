// ertgpu_decode_file -- decode a raw uint8 IQ capture (rtl_sdr / rtlamr -samplefile format) or stdin
// through the C++ mirror of rtlamr's receive loop (receiver.hpp) on the GPU and print one line per message
// in rtlamr's plain format without the time fields (protocol/parse.go:119-121 StringNoOffset).
//   ertgpu_decode_file [-msgtype=scm,idm|all] [-symbollength=72] [-blockdedup=true] [-blocks=4096] [-device=0]
//                      [-filterid=1,2] [-filtertype=4,5] [-unique=true] [-format=plain|csv] FILE|- [FILE ...]
// -filterid / -filtertype / -unique and -format=csv are rtlamr's flags of the same names (flags.go:69-70,51,140-151; the filter
// chain runs before the cross-block dedup, main.go:236-260).  With -format=csv a line is LogMessage.Record() (parse.go:123-129:
// time, offset, length, then the message's own record) through Go's encoding/csv rules.
// Several files are a batch of independent streams through ONE decoder (buffers, tables and kernels set up once):
// each starts from zeroed history with block numbers from 0, exactly as if it were decoded alone.
#include <cstdio>
#include <cstring>
#include <ctime>
#include <iostream>
#include <vector>

#include "receiver.hpp"

int main(int argc, char** argv) {
    std::string msgtype = "scm";
    std::vector<std::string> files;
    int chip = 72, device = 0;
    long long blocks = 4096;
    bool block_dedup = true, quiet = false, unique = false, csv = false;
    std::string filterid, filtertype;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a.rfind("-msgtype=", 0) == 0) msgtype = a.substr(9);
        else if (a.rfind("-symbollength=", 0) == 0) chip = atoi(a.c_str() + 14);
        else if (a.rfind("-device=", 0) == 0) device = atoi(a.c_str() + 8);
        else if (a.rfind("-blocks=", 0) == 0) blocks = atoll(a.c_str() + 8);
        else if (a.rfind("-blockdedup=", 0) == 0) block_dedup = a.substr(12) != "false";
        else if (a.rfind("-filterid=", 0) == 0) filterid = a.substr(10);
        else if (a.rfind("-filtertype=", 0) == 0) filtertype = a.substr(12);
        else if (a.rfind("-unique=", 0) == 0) unique = a.substr(8) == "true";
        else if (a == "-unique") unique = true;
        else if (a.rfind("-format=", 0) == 0) csv = a.substr(8) == "csv";
        else if (a == "-quiet") quiet = true;
        else files.push_back(a);
    }
    if (files.empty()) {
        fprintf(stderr, "usage: %s [-msgtype=scm,scm+,idm,netidm,r900,r900bcd|all] [-symbollength=N] [-blockdedup=true|false] "
                        "[-blocks=K] [-device=D] [-filterid=ID,..] [-filtertype=T,..] [-unique=true|false] [-format=plain|csv] [-quiet] FILE|- [FILE ...]\n", argv[0]);
        return 2;
    }
    try {
        receiver::Receiver rcvr(msgtype, chip, device, blocks);
        fputs(rcvr.decoder().Log().c_str(), stderr);
        // the order flag.Visit gives the chain (main.go:97-113): filterid, filtertype, unique
        if (!filterid.empty()) {
            auto f = std::make_shared<receiver::MeterIDFilter>();
            f->ids.Set(filterid);
            rcvr.filters().Add(f);
        }
        if (!filtertype.empty()) {
            auto f = std::make_shared<receiver::MeterTypeFilter>();
            f->types.Set(filtertype);
            rcvr.filters().Add(f);
        }
        if (unique) rcvr.filters().Add(std::make_shared<receiver::UniqueFilter>());
        for (size_t fi = 0; fi < files.size(); fi++) {
            const std::string& file = files[fi];
            FILE* in = file == "-" ? stdin : fopen(file.c_str(), "rb");
            if (!in) {
                perror(file.c_str());
                return 1;
            }
            if (fi) rcvr.Reset();
            auto st = rcvr.Run(in, block_dedup, [&](const protocol::Message& m) {
                if (quiet) return;
                if (csv) {
                    receiver::LogMessage lm;
                    timespec ts{};
                    clock_gettime(CLOCK_REALTIME, &ts);
                    lm.unix_seconds = (int64_t)ts.tv_sec;
                    lm.nanos = (int32_t)ts.tv_nsec;
                    lm.Message = &m;
                    fputs(receiver::CsvLine(lm.Record()).c_str(), stdout);
                } else {
                    printf("{Block:%lld Idx:%d %s:%s}\n", (long long)m.Block, m.Idx, m.MsgType().c_str(), m.String().c_str());
                }
            });
            if (in != stdin) fclose(in);
            fprintf(stderr, "%s%s%lld messages (%lld duplicates dropped) from %lld blocks, %.1f MB in %.3f s = %.1f Msamples/s\n",
                    files.size() > 1 ? file.c_str() : "", files.size() > 1 ? ": " : "", (long long)st.messages,
                    (long long)st.duplicates, (long long)st.blocks, st.bytes / 1e6, st.seconds,
                    st.bytes / 2e6 / (st.seconds > 0 ? st.seconds : 1));
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
