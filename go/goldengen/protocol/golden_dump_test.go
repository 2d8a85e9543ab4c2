// golden_dump_test.go -- dumps the STOCK protocol.Decoder's internal state as golden vectors.
//
// Drop this file into <rtlamr>/protocol/ (package protocol: it reads the unexported csum / packed
// buffers) and run, on any machine with Go, from the rtlamr checkout:
//
//	GOLDEN_IN=$PWD/assets/sample.bin GOLDEN_OUT=/tmp/go_dump_sample_cl78_scm.jsonl \
//	GOLDEN_MSGTYPES=scm GOLDEN_CL=78 go test ./protocol -run TestDumpGolden -count=1
//
// (GOLDEN_IN must be absolute: `go test` runs the test binary inside ./protocol.  tests/golden/make_go_golden.sh in the
// rtlamr_b200 repository runs every case the parity loader knows about).  The output is JSON lines:
//
//	{"kind":"config", ...PacketConfig after Allocate...}
//	{"kind":"cand","block":B,"preamble":"1111...","idx":I,"bytes":"hex"}      every Data of every Decode call
//	{"kind":"tap","block":B,"signal":"hex f32le","csum":"hex f32le","quantized":"hex","packed":"hex"}
//
// Taps are written for the blocks listed in GOLDEN_TAP_BLOCKS (comma separated; default: the first three
// blocks, every block that produced a candidate, and the last block).  `packed` is the state after the LAST
// Search of the block (Go iterates its preamble map in random order; with one preamble it is deterministic,
// and Pack does not depend on the preamble anyway: decode.go:259-265).
//
// No parser package is imported (that would be an import cycle inside package protocol): the parsers'
// PacketConfig literals are restated below (scm/scm.go:42-50, scmplus/scmplus.go:49-57, idm/idm.go:48-56,
// netidm/netidm.go:60-68, r900/r900.go:57-65) and a recording Parser stands in for them -- it receives
// exactly the []Data the real parsers receive (decode.go:179-187).
package protocol

import (
	"bufio"
	"encoding/binary"
	"encoding/hex"
	"encoding/json"
	"math"
	"os"
	"sort"
	"strconv"
	"strings"
	"sync"
	"testing"
)

type goldenCand struct {
	Kind     string `json:"kind"`
	Block    int    `json:"block"`
	Preamble string `json:"preamble"`
	Idx      int    `json:"idx"`
	Bytes    string `json:"bytes"`
}

type recordingParser struct {
	cfg   PacketConfig
	mu    *sync.Mutex
	block *int
	out   *[]goldenCand
}

func (p recordingParser) SetDecoder(*Decoder) {}
func (p recordingParser) Cfg() PacketConfig   { return p.cfg }
func (p recordingParser) Parse(pkts []Data, msgCh chan Message, wg *sync.WaitGroup) {
	p.mu.Lock()
	for _, pkt := range pkts {
		*p.out = append(*p.out, goldenCand{"cand", *p.block, p.cfg.Preamble, pkt.Idx, hex.EncodeToString(pkt.Bytes)})
	}
	p.mu.Unlock()
	wg.Done()
}

func goldenConfig(name string, chipLength int) (PacketConfig, bool) {
	base := PacketConfig{Protocol: name, CenterFreq: 912600155, DataRate: 32768, ChipLength: chipLength}
	switch name {
	case "scm":
		base.Preamble, base.PreambleSymbols, base.PacketSymbols = "111110010101001100000", 21, 96
	case "scm+":
		base.Preamble, base.PreambleSymbols, base.PacketSymbols = "0001011010100011", 16, 128
	case "idm", "netidm":
		base.Preamble, base.PreambleSymbols, base.PacketSymbols = "01010101010101010001011010100011", 32, 92*8
	case "r900", "r900bcd":
		base.Preamble, base.PreambleSymbols, base.PacketSymbols = "00000000000000001110010101100100", 32, 116
		base.CenterFreq = 912380000
	default:
		return base, false
	}
	return base, true
}

func f32hex(v []float32) string {
	b := make([]byte, 4*len(v))
	for i, x := range v {
		binary.LittleEndian.PutUint32(b[4*i:], math.Float32bits(x))
	}
	return hex.EncodeToString(b)
}

func TestDumpGolden(t *testing.T) {
	in, out := os.Getenv("GOLDEN_IN"), os.Getenv("GOLDEN_OUT")
	if in == "" || out == "" {
		t.Skip("GOLDEN_IN / GOLDEN_OUT not set")
	}
	msgtypes := strings.Split(os.Getenv("GOLDEN_MSGTYPES"), ",")
	cl, err := strconv.Atoi(os.Getenv("GOLDEN_CL"))
	if err != nil {
		t.Fatal("GOLDEN_CL:", err)
	}
	iq, err := os.ReadFile(in)
	if err != nil {
		t.Fatal(err)
	}

	d := NewDecoder()
	var mu sync.Mutex
	var cands []goldenCand
	block := 0
	seenPre := map[string]bool{}
	for _, name := range msgtypes {
		cfg, ok := goldenConfig(strings.TrimSpace(name), cl)
		if !ok {
			t.Fatalf("unknown msgtype %q", name)
		}
		// one recorder per DISTINCT preamble is enough for the candidate list; the others only merge their config
		rec := recordingParser{cfg, &mu, &block, &cands}
		if seenPre[cfg.Preamble] {
			sink := []goldenCand{}
			rec.out = &sink
		}
		seenPre[cfg.Preamble] = true
		d.RegisterProtocol(rec)
	}
	d.Allocate()

	f, err := os.Create(out)
	if err != nil {
		t.Fatal(err)
	}
	defer f.Close()
	w := bufio.NewWriter(f)
	defer w.Flush()
	enc := json.NewEncoder(w)
	enc.Encode(map[string]interface{}{"kind": "config", "msgtypes": msgtypes, "chip_length": cl, "cfg": d.Cfg,
		"input": in, "input_bytes": len(iq)})

	nblocks := len(iq) / d.Cfg.BlockSize2
	want := map[int]bool{0: true, 1: true, 2: true, nblocks - 1: true}
	explicit := false
	if s := os.Getenv("GOLDEN_TAP_BLOCKS"); s != "" {
		explicit = true
		want = map[int]bool{}
		for _, x := range strings.Split(s, ",") {
			if b, err := strconv.Atoi(strings.TrimSpace(x)); err == nil {
				want[b] = true
			}
		}
	}
	for block = 0; block < nblocks; block++ {
		before := len(cands)
		for range d.Decode(iq[block*d.Cfg.BlockSize2 : (block+1)*d.Cfg.BlockSize2]) { // main.go:235
		}
		if want[block] || (!explicit && len(cands) > before) {
			enc.Encode(map[string]interface{}{"kind": "tap", "block": block, "signal": f32hex(d.Signal),
				"csum": f32hex(d.csum), "quantized": hex.EncodeToString(d.Quantized), "packed": hex.EncodeToString(d.packed)})
		}
	}
	sort.Slice(cands, func(i, j int) bool {
		a, b := cands[i], cands[j]
		if a.Block != b.Block {
			return a.Block < b.Block
		}
		if a.Preamble != b.Preamble {
			return a.Preamble < b.Preamble
		}
		return a.Idx < b.Idx
	})
	for _, c := range cands {
		enc.Encode(c)
	}
	enc.Encode(map[string]interface{}{"kind": "end", "blocks": nblocks, "candidates": len(cands)})
}
