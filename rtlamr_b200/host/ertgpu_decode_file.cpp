// ertgpu_decode_file -- decode a raw uint8 IQ capture (rtl_sdr / rtlamr -samplefile format) or stdin
// through the C++ mirror of rtlamr's receive loop (receiver.hpp) on the GPU and print one line per message
// in rtlamr's plain format without the time fields (protocol/parse.go:119-121 StringNoOffset).
//   ertgpu_decode_file [-msgtype=scm,idm|all] [-symbollength=72] [-blockdedup=true] [-blocks=4096] [-device=0] FILE|- [FILE ...]
// Several files are a batch of independent streams through ONE decoder (buffers, tables and kernels set up once):
// each starts from zeroed history with block numbers from 0, exactly as if it were decoded alone.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <vector>

#include "receiver.hpp"

int main(int argc, char** argv) {
    std::string msgtype = "scm";
    std::vector<std::string> files;
    int chip = 72, device = 0;
    long long blocks = 4096;
    bool block_dedup = true, quiet = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a.rfind("-msgtype=", 0) == 0) msgtype = a.substr(9);
        else if (a.rfind("-symbollength=", 0) == 0) chip = atoi(a.c_str() + 14);
        else if (a.rfind("-device=", 0) == 0) device = atoi(a.c_str() + 8);
        else if (a.rfind("-blocks=", 0) == 0) blocks = atoll(a.c_str() + 8);
        else if (a.rfind("-blockdedup=", 0) == 0) block_dedup = a.substr(12) != "false";
        else if (a == "-quiet") quiet = true;
        else files.push_back(a);
    }
    if (files.empty()) {
        fprintf(stderr, "usage: %s [-msgtype=scm,scm+,idm,netidm,r900,r900bcd|all] [-symbollength=N] [-blockdedup=true|false] "
                        "[-blocks=K] [-device=D] [-quiet] FILE|- [FILE ...]\n", argv[0]);
        return 2;
    }
    try {
        receiver::Receiver rcvr(msgtype, chip, device, blocks);
        fputs(rcvr.decoder().Log().c_str(), stderr);
        for (size_t fi = 0; fi < files.size(); fi++) {
            const std::string& file = files[fi];
            FILE* in = file == "-" ? stdin : fopen(file.c_str(), "rb");
            if (!in) {
                perror(file.c_str());
                return 1;
            }
            if (fi) rcvr.Reset();
            auto st = rcvr.Run(in, block_dedup, [&](const protocol::Message& m) {
                if (!quiet) printf("{Block:%lld Idx:%d %s:%s}\n", (long long)m.Block, m.Idx, m.MsgType().c_str(), m.String().c_str());
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
