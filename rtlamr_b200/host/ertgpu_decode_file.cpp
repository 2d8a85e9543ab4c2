// ertgpu_decode_file -- decode a raw uint8 IQ capture (rtl_sdr / rtlamr -samplefile format) through the
// C++ mirror of protocol.Decoder + parsers on the GPU and print one line per message in rtlamr's plain
// format without the time fields (protocol/parse.go:119-121 StringNoOffset).
//   ertgpu_decode_file [-msgtype=scm,idm] [-symbollength=72] [-device=0] FILE
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

#include "protocol.hpp"

int main(int argc, char** argv) {
    std::string msgtype = "scm", file;
    int chip = 72, device = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a.rfind("-msgtype=", 0) == 0) msgtype = a.substr(9);
        else if (a.rfind("-symbollength=", 0) == 0) chip = atoi(a.c_str() + 14);
        else if (a.rfind("-device=", 0) == 0) device = atoi(a.c_str() + 8);
        else file = a;
    }
    if (file.empty()) {
        fprintf(stderr, "usage: %s [-msgtype=scm,scm+,idm,netidm,r900,r900bcd|all] [-symbollength=N] FILE\n", argv[0]);
        return 2;
    }
    if (msgtype == "all") msgtype = "scm,scm+,idm,r900";  // main.go:67-73
    try {
        protocol::RegisterStockParsers();
        protocol::Decoder d;
        std::stringstream ss(msgtype);
        std::string name;
        while (std::getline(ss, name, ',')) d.RegisterProtocol(protocol::NewParser(name, chip));
        d.Allocate(device);
        fputs(d.Log().c_str(), stderr);
        std::ifstream f(file, std::ios::binary);
        std::vector<uint8_t> iq((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        const size_t whole = iq.size() / (size_t)d.Cfg.BlockSize2 * (size_t)d.Cfg.BlockSize2;  // main.go:166-186 reads whole blocks only
        auto msgs = d.Decode(iq.data(), whole);
        for (auto& m : msgs) printf("{Block:%lld Idx:%d %s:%s}\n", (long long)m->Block, m->Idx, m->MsgType().c_str(), m->String().c_str());
        fprintf(stderr, "%zu messages from %zu blocks\n", msgs.size(), whole / (size_t)d.Cfg.BlockSize2);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
