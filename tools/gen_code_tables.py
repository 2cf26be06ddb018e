#!/usr/bin/env python3
"""Generate gr-dvbs2rx_b200/csrc/dvbs2_code_tables.inc from the reference checkout.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tools/gen_code_tables.py [/root/reference]

What it reads (numeric data of ETSI EN 302 307-1 / -2 and EN 302 755, nothing else):
  * lib/dvb_s2_tables.hh, lib/dvb_s2x_tables.hh, lib/dvb_t2_tables.hh
      -> per code: N, K and the parity-accumulator address rows (DEG/LEN/POS).
  * lib/ldpc_decoder_bb_impl.cc:104-307 -> (framesize, rate, standard) -> table name.
  * lib/fec_params.cc:16-344            -> (framesize, rate) -> kbch, nbch, t.
  * include/gnuradio/dvbs2rx/dvb_config.h:15-121 -> enum ordinals.

What it writes is NOT the reference's layout.  Each code is re-expressed as the list of
circulants the B200 kernels consume: for accumulator address x = q*a + i of 360-bit group g
(lib/ldpc_decoder/ldpc.hh:67-78) check node (layer i, lane j) reads bit
g*360 + ((j - a) mod 360).  One uint32 per circulant, sorted by (layer, group, shift):

    word = layer << 17 | group << 9 | shift        (layer < 136, group < 180, shift < 360)
"""
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "gr-dvbs2rx_b200", "csrc", "dvbs2_code_tables.inc")


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def parse_enum(text, name):
    m = re.search(r"enum\s+%s\s*\{(.*?)\}" % name, text, re.S)
    names = []
    val = 0
    for item in m.group(1).split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            k, v = item.split("=")
            item, val = k.strip(), int(v.strip(), 0)
        names.append((item, val))
        val += 1
    return dict(names)


def parse_tables(text):
    tables = {}
    for m in re.finditer(r"struct\s+(DVB_\w+)\s*\{(.*?)\n\};", text, re.S):
        name, body = m.group(1), m.group(2)

        def scalar(key):
            return int(re.search(r"\b%s\s*=\s*(\d+)" % key, body).group(1))

        def array(key):
            mm = re.search(r"\b%s\[\]\s*=\s*\{(.*?)\}" % key, body, re.S)
            return [int(x) for x in re.findall(r"\d+", mm.group(1))]

        tables[name] = dict(M=scalar("M"), N=scalar("N"), K=scalar("K"),
                            LINKS_MAX_CN=scalar("LINKS_MAX_CN"),
                            LINKS_TOTAL=scalar("LINKS_TOTAL"),
                            DEG=array("DEG"), LEN=array("LEN"), POS=array("POS"))
    return tables


def circulants(t):
    """(layer, group, shift) triples of one code, following LDPC<TABLE>::next_group."""
    M, N, K = t["M"], t["N"], t["K"]
    assert M == 360
    R = N - K
    q = R // M
    assert q * M == R
    out = []
    pos = t["POS"]
    p = 0
    g = 0
    for deg, ln in zip(t["DEG"], t["LEN"]):
        if ln == 0:
            break
        for _ in range(ln):
            row = pos[p:p + deg]
            p += deg
            for x in row:
                assert 0 <= x < R
                out.append((x % q, g, x // q))
            g += 1
    assert g * M == K, (g, K)
    assert p == len(pos), (p, len(pos))
    out.sort()
    # data links + 2 parity links per check, minus the missing link of check (0,0)
    assert len(out) * M + 2 * R - 1 == t["LINKS_TOTAL"], (len(out), t["LINKS_TOTAL"])
    return q, out


def parse_rate_map(text):
    """(framesize, rate) -> [(standard or None, table)] from the constructor's switch."""
    sect = text[text.index("if (framesize == FECFRAME_NORMAL)"):text.index("decode = nullptr")]
    fs = None
    rates = []
    std = None
    res = []
    for line in sect.splitlines():
        line = line.strip()
        m = re.search(r"framesize == (FECFRAME_\w+)", line)
        if m:
            fs = m.group(1)
            continue
        if line.startswith("} else {") and fs == "FECFRAME_SHORT" and not rates:
            fs = "FECFRAME_MEDIUM"
            continue
        m = re.match(r"case (\w+):", line)
        if m:
            rates.append(m.group(1))
            std = None
            continue
        if "standard == STANDARD_DVBS2" in line:
            std = "STANDARD_DVBS2"
            continue
        if line.startswith("} else {") and std == "STANDARD_DVBS2":
            std = "STANDARD_DVBT2"
            continue
        m = re.search(r"new LDPC<(\w+)>", line)
        if m:
            for r in rates:
                res.append((fs, r, std, m.group(1)))
            continue
        if line.startswith("break;"):
            rates = []
            std = None
    return res


def parse_fec_params(text):
    res = {}
    fs = None
    rates = []
    cur = {}
    for line in text.splitlines():
        line = line.strip()
        m = re.search(r"framesize == (FECFRAME_\w+)", line)
        if m:
            fs = m.group(1)
            continue
        if line.startswith("} else {") and fs == "FECFRAME_SHORT":
            fs = "FECFRAME_MEDIUM"
            continue
        m = re.match(r"case (\w+):", line)
        if m:
            rates.append(m.group(1))
            continue
        m = re.match(r"fec_info\.bch\.(\w) = (\d+);", line)
        if m:
            cur[m.group(1)] = int(m.group(2))
            continue
        if line.startswith("break;"):
            if cur:
                for r in rates:
                    res[(fs, r)] = (cur["k"], cur["n"], cur["t"])
            rates, cur = [], {}
    return res


def main():
    cfg = read("include/gnuradio/dvbs2rx/dvb_config.h")
    rate_enum = parse_enum(cfg, "dvb_code_rate_t")
    fs_enum = parse_enum(cfg, "dvb_framesize_t")
    std_enum = parse_enum(cfg, "dvb_standard_t")

    tables = {}
    for f in ("lib/dvb_s2_tables.hh", "lib/dvb_s2x_tables.hh", "lib/dvb_t2_tables.hh"):
        tables.update(parse_tables(read(f)))
    assert len(tables) == 57, len(tables)

    rmap = parse_rate_map(read("lib/ldpc_decoder_bb_impl.cc"))
    fec = parse_fec_params(read("lib/fec_params.cc"))

    names = sorted(tables, key=lambda n: (n.split("_TABLE_")[0], n.split("_TABLE_")[1][0],
                                          int(n.split("_TABLE_")[1][1:])))
    index = {n: i for i, n in enumerate(names)}

    L = []
    L.append("// GENERATED by tools/gen_code_tables.py -- do not edit.")
    L.append("// DVB-S2 / S2X / T2 LDPC parity-address data (ETSI EN 302 307-1 Annex B/C,")
    L.append("// EN 302 307-2 Annex B/C, EN 302 755 Annex A/B) re-expressed as circulants:")
    L.append("//   word = layer << 17 | group << 9 | shift ; check (layer, j) reads data bit")
    L.append("//   group*360 + ((j - shift) mod 360).  Sorted by (layer, group, shift).")
    L.append("// Cross-reference: reference lib/dvb_s2_tables.hh, dvb_s2x_tables.hh,")
    L.append("// dvb_t2_tables.hh (accumulator address x = q*shift + layer of that group).")
    L.append("")
    for n in names:
        q, tr = circulants(tables[n])
        t = tables[n]
        tables[n]["q"] = q
        tables[n]["ntr"] = len(tr)
        L.append("static const uint32_t kCirc_%s[%d] = {" % (n, len(tr)))
        words = ["0x%07x" % ((l << 17) | (g << 9) | a) for (l, g, a) in tr]
        for k in range(0, len(words), 8):
            L.append("    " + ", ".join(words[k:k + 8]) + ",")
        L.append("};")
    L.append("")
    L.append("static const Dvbs2LdpcTableDef kLdpcTables[%d] = {" % len(names))
    L.append("    // name, N, K, q, n_circulants, links_total, links_max_cn, circulants")
    for n in names:
        t = tables[n]
        L.append('    { "%s", %d, %d, %d, %d, %d, %d, kCirc_%s },'
                 % (n, t["N"], t["K"], t["q"], t["ntr"], t["LINKS_TOTAL"], t["LINKS_MAX_CN"], n))
    L.append("};")
    L.append("")
    L.append("// (framesize, rate, standard) -> LDPC table and BCH parameters.")
    L.append("// standard = -1: any.  Ordinals follow include/gnuradio/dvbs2rx/dvb_config.h:15-121.")
    L.append("// kbch/nbch/t follow lib/fec_params.cc:16-344 (0 when the reference gives none).")
    rows = []
    for fs, r, std, tab in rmap:
        k, n, t = fec.get((fs, r), (0, 0, 0))
        rows.append("    { %d, %d, %d, %d, %d, %d, %d }, // %s %s %s -> %s"
                    % (fs_enum[fs], rate_enum[r], -1 if std is None else std_enum[std],
                       index[tab], k, n, t, fs, r, std or "", tab))
    L.append("static const Dvbs2ModcodDef kModcods[%d] = {" % len(rows))
    L.append("    // framesize, rate, standard, table, kbch, nbch, t")
    L.extend(rows)
    L.append("};")
    L.append("")
    with open(OUT, "w") as f:
        f.write("\n".join(L))
    print("wrote %s: %d tables, %d modcod rows" % (os.path.normpath(OUT), len(names), len(rows)))
    mx = max(t["LINKS_MAX_CN"] for t in tables.values())
    print("max CN degree", mx, "max q", max(t["q"] for t in tables.values()),
          "max circulants", max(t["ntr"] for t in tables.values()))


if __name__ == "__main__":
    main()
