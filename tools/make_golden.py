"""Extracts the reference's own known-answer data for the hot path into tests/golden/*.json.

Run HERE (where /root/reference exists):  python tools/make_golden.py
Only DATA (test vectors, expected digests) is extracted -- no reference code.  Where a vector needs a
reference-side preprocessing step that is out of scope for the GPU engine (DER parsing of Wycheproof
signatures), the reference compiled into oracle/_ref does that step once and the result is stored.

Sources (paths relative to /root/reference):
  src/modules/schnorrsig/tests_impl.h:210-767    BIP-340 vectors 0..18
  src/tests.c:4617-4674                          run_ecmult_chain parameters + endpoint
  src/tests.c:5800-5960                          run_ecmult_constants digests
  src/wycheproof/ecdsa_secp256k1_sha256_bitcoin_test.h   463 ECDSA vectors (driver src/tests.c:7803)
  src/modules/rangeproof/tests_impl.h:589-812, :883-1349  rangeproof vectors
  src/modules/generator/tests_impl.h             pedersen / generator fixed vectors
  src/modules/bppp/test_vectors/verify.h         BP++ norm-argument verify vectors
  src/modules/recovery/tests_impl.h:156-249      ECDSA recovery edge cases
  src/modules/schnorrsig_halfagg/tests_impl.h:72-167   half-aggregation spec vectors
"""
import ctypes
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

ARRAY_RE = re.compile(r"(?:static\s+)?(?:const\s+)?unsigned\s+char\s+(\w+)\s*\[[^\]]*\]\s*=\s*\{([^}]*)\}\s*;", re.S)


def parse_arrays(text):
    """[(name, bytes, start_offset)] for every `unsigned char NAME[..] = { 0x.., ... };` in text."""
    out = []
    for m in ARRAY_RE.finditer(text):
        body = m.group(2)
        vals = re.findall(r"0x([0-9a-fA-F]{1,2})|\b(\d+)\b", body)
        bs = bytes(int(h, 16) if h else int(d) for h, d in vals)
        out.append((m.group(1), bs, m.start()))
    return out


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def func_body(text, name):
    i = text.index(name + "(void) {")
    depth = 0
    j = text.index("{", i)
    k = j
    while True:
        if text[k] == "{":
            depth += 1
        elif text[k] == "}":
            depth -= 1
            if depth == 0:
                return text[j:k + 1]
        k += 1


def bip340():
    body = func_body(read("src/modules/schnorrsig/tests_impl.h"), "static void test_schnorrsig_bip_vectors")
    parts = re.split(r"/\* Test vector (\d+) \*/", body)
    vectors = []
    for idx, blk in zip(parts[1::2], parts[2::2]):
        arrs = {n: b for n, b, _ in parse_arrays(blk)}
        v = {"index": int(idx), "pk": arrs["pk"].hex()}
        if "CHECK(!secp256k1_xonly_pubkey_parse" in blk:
            v["pk_parse_fails"] = True
            v["expected"] = 0
        else:
            m = re.search(r"check_verify\(pk,\s*(\w+),\s*([^,]+),\s*sig,\s*(\d)\)", blk)
            v["pk_parse_fails"] = False
            v["expected"] = int(m.group(3))
            if m.group(1) == "NULL":
                v["msg"] = ""
            elif "msg" in arrs:
                v["msg"] = arrs["msg"].hex()
            else:
                # vector 18: `unsigned char msg[100]; memset(msg, 0x99, sizeof(msg));`
                mm = re.search(r"unsigned char msg\[(\d+)\];\s*memset\(msg, 0x([0-9a-fA-F]+), sizeof\(msg\)\);", blk)
                v["msg"] = (bytes([int(mm.group(2), 16)]) * int(mm.group(1))).hex()
            v["sig"] = arrs["sig"].hex()
        vectors.append(v)
    assert [v["index"] for v in vectors] == list(range(19)), [v["index"] for v in vectors]
    return {"source": "src/modules/schnorrsig/tests_impl.h:210-767 (BIP-340 test-vectors.csv 0..18)", "vectors": vectors}


def words_hex(s):
    return "".join("%08x" % int(w, 16) for w in re.findall(r"0x([0-9a-fA-F]+)", s))


def ecmult_golden():
    t = read("src/tests.c")
    chain = t[t.index("static void run_ecmult_chain(void)"):t.index("static void test_point_times_order")]
    consts = re.findall(r"SECP256K1_(?:GEJ|SCALAR)_CONST\(([^)]*)\)", chain)
    a, xn, gn, xf, gf, rp = consts[0], consts[1], consts[2], consts[3], consts[4], consts[5]
    out = {
        "source": "src/tests.c:4617-4674 (run_ecmult_chain), :5800-5960 (run_ecmult_constants)",
        "chain": {"a_xy": words_hex(a), "xn": words_hex(xn), "gn": words_hex(gn), "xf": "1337", "gf": "7113",
                  "iterations": 20000, "expected_xy": words_hex(rp)},
    }
    assert len(out["chain"]["a_xy"]) == 128 and len(out["chain"]["expected_xy"]) == 128
    seg = t[t.index("static void test_ecmult_constants_2bit(void)"):t.index("static void test_ecmult_gen_blind")]
    arrs = {n: b for n, b, _ in parse_arrays(seg)}
    out["constants"] = {
        "2bit": arrs["expected32"].hex(),
        "sha_6bit20": {"prefix": 4808378, "iter": 1024, "digest": arrs["expected32_6bit20"].hex()},
        "sha_8bit8": {"prefix": 1607366309, "iter": 2048, "digest": arrs["expected32_8bit8"].hex()},
    }
    assert "test_ecmult_constants_sha(4808378u, 1024, expected32_6bit20)" in seg
    assert "test_ecmult_constants_sha(1607366309u, 2048, expected32_8bit8)" in seg
    return out


def wycheproof(ref):
    t = read("src/wycheproof/ecdsa_secp256k1_sha256_bitcoin_test.h")
    arrs = {n: b for n, b, _ in parse_arrays(t)}
    msgs, pks, sigs = arrs["wycheproof_ecdsa_messages"], arrs["wycheproof_ecdsa_public_keys"], arrs["wycheproof_ecdsa_signatures"]
    tv = t[t.index("testvectors[SECP256K1_ECDSA_WYCHEPROOF_NUMBER_TESTVECTORS]"):]
    rows = re.findall(r"/\* tcId: (\d+)\.\s*(.*?) \*/\s*\{(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\s*\}", tv)
    assert len(rows) == 463, len(rows)
    L = ref.lib
    vectors = []
    for tc, comment, pko, mo, ml, so, sl, exp in rows:
        pko, mo, ml, so, sl, exp = map(int, (pko, mo, ml, so, sl, exp))
        pk65 = pks[pko:pko + 65]
        msg = msgs[mo:mo + ml]
        der = sigs[so:so + sl]
        sig = ctypes.create_string_buffer(64)
        der_ok = L.secp256k1_ecdsa_signature_parse_der(ref.ctx, sig, der, ctypes.c_size_t(sl))
        compact = ctypes.create_string_buffer(64)
        if der_ok:
            L.secp256k1_ecdsa_signature_serialize_compact(ref.ctx, compact, sig)
        vectors.append({"tcId": int(tc), "comment": comment, "pk65": pk65.hex(), "msghash": hashlib.sha256(msg).hexdigest(),
                        "der": der.hex(), "der_ok": int(der_ok), "compact": compact.raw.hex() if der_ok else "", "expected": exp})
    return {"source": "src/wycheproof/ecdsa_secp256k1_sha256_bitcoin_test.h (driver src/tests.c:7803-7831); DER parsed by the reference",
            "vectors": vectors}


def rangeproof(ref):
    t = read("src/modules/rangeproof/tests_impl.h")
    L = ref.lib
    gen_h = ctypes.c_void_p.in_dll(L, "secp256k1_generator_h") if False else None
    out = []
    for fn, tag in (("static void test_rangeproof_fixed_vectors", "fixed"), ("static void test_rangeproof_fixed_vectors_reproducible", "reproducible")):
        body = func_body(t, fn)
        arrs = {n: b for n, b, _ in parse_arrays(body)}
        for name in sorted(a for a in arrs if a.startswith("vector_")):
            k = name.split("_")[1]
            proof, commit = arrs[name], arrs["commit_" + k]
            ok, mn, mx = ref.rangeproof_verify_one(commit, proof)
            assert ok == 1, (tag, name)
            out.append({"name": "%s_%s" % (tag, k), "proof": proof.hex(), "commit": commit.hex(), "expected": 1,
                        "min_value": mn, "max_value": mx})
    # the expectations spelled out in the reference test text
    chk = {v["name"]: (v["min_value"], v["max_value"]) for v in out}
    assert chk["fixed_1"] == (86, 25586) and chk["fixed_2"] == (0, 15) and chk["fixed_3"] == (2**64 - 1, 2**64 - 1)
    assert chk["reproducible_0"] == (0, 2**64 - 1) and chk["reproducible_1"] == (3, 73)
    return {"source": "src/modules/rangeproof/tests_impl.h:589-812 and :883-1349; generator = secp256k1_generator_h", "vectors": out}


def bppp_vectors(ref):
    """verify_vector_0..12 (src/modules/bppp/test_vectors/verify.h; driver tests_impl.h:540-598: fresh SHA-256 transcript).
    Generators / commitments are stored both as in the file (33-byte compressed) and decompressed by the reference (x||y)."""
    t = read("src/modules/bppp/test_vectors/verify.h")
    arrs = {n: b for n, b, _ in parse_arrays(t)}
    L = ref.lib

    def decompress(c33):
        if c33 == bytes(33):
            return bytes(64)
        pk = ctypes.create_string_buffer(64)
        assert L.secp256k1_ec_pubkey_parse(ref.ctx, pk, c33, ctypes.c_size_t(33)) == 1
        xy = ctypes.create_string_buffer(64)
        L.ref_pubkey_to_xy(xy, pk)
        return xy.raw
    gens33 = arrs["verify_vector_gens"]
    gens_xy = b"".join(decompress(gens33[33 * i:33 * i + 33]) for i in range(len(gens33) // 33))
    out = []
    for i in range(13):
        pre = "verify_vector_%d_" % i
        g_len = int(re.search(pre + r"n_vec_len = (\d+);", t).group(1))
        c_raw = re.search(pre + r"c_vec32\[(\d+)\]\[32\] = \{(.*?)\};", t, re.S)
        h_len = int(c_raw.group(1))
        cv = bytes(int(h, 16) for h in re.findall(r"0x([0-9a-fA-F]{2})", c_raw.group(2)))
        assert len(cv) == 32 * h_len
        result = int(re.search(pre + r"result = (\d);", t).group(1))
        out.append({"index": i, "g_len": g_len, "h_len": h_len, "proof": arrs[pre + "proof"].hex(), "r32": arrs[pre + "r32"].hex(),
                    "c_vec32": cv.hex(), "commit33": arrs[pre + "commit33"].hex(), "commit_xy": decompress(arrs[pre + "commit33"]).hex(),
                    "expected": result})
    return {"source": "src/modules/bppp/test_vectors/verify.h (driver src/modules/bppp/tests_impl.h:540-598, plain SHA-256 transcript)",
            "gens33": gens33.hex(), "gens_xy": gens_xy.hex(), "vectors": out}


def rewind_vectors():
    """src/modules/rangeproof/tests_impl.h:843-851, :883-1352 (test_rangeproof_fixed_vectors_reproducible): what rewinding the three
    reproducible vectors with vector_nonce must return (blind = vector_blind, the value, an all-0xFF message of the stated length)."""
    text = read("src/modules/rangeproof/tests_impl.h")
    arr = {n: b for n, b, _ in parse_arrays(text)}
    gold = json.load(open(os.path.join(OUT, "rangeproof.json")))["vectors"]
    by = {v["name"]: v for v in gold}
    exp = [("reproducible_0", 2**64 - 1, 3968), ("reproducible_1", 13, 128), ("reproducible_2", 2**63 - 1, 0)]
    m = re.search(r"#define\s+SECP256K1_RANGEPROOF_MAX_MESSAGE_LEN\s+(\d+)", read("include/secp256k1_rangeproof.h"))
    assert m and int(m.group(1)) == 3968
    out = []
    for name, value, mlen in exp:
        v = by[name]
        out.append({"name": name, "commit": v["commit"], "proof": v["proof"], "nonce": arr["vector_nonce"].hex(), "blind": arr["vector_blind"].hex(),
                    "value": str(value), "message_len": mlen, "message_byte": "ff", "min_value": v["min_value"], "max_value": v["max_value"]})
    return {"source": "src/modules/rangeproof/tests_impl.h:843-881, :883-1352", "vectors": out}


def recovery_vectors():
    """src/modules/recovery/tests_impl.h:156-199 (test_ecdsa_recovery_edge_cases): which recids recover a key."""
    text = read("src/modules/recovery/tests_impl.h")
    body = re.sub(r"/\*.*?\*/", "", func_body(text, "test_ecdsa_recovery_edge_cases"), flags=re.S)
    body = re.sub(r"'(.)'", lambda m: "0x%02x" % ord(m.group(1)), body)      # msg32 is spelled with character literals
    arr = {n: b for n, b, _ in parse_arrays(body)}
    assert arr["msg32"] == b"This is a very secret message..." and len(arr["sig64"]) == 64 and len(arr["sigb64"]) == 64
    vec = [{"msg32": arr["msg32"].hex(), "sig64": arr["sig64"].hex(), "recover_ok": [0, 1, 0, 0],
            "note": "CHECK(!recover) for recid 0, 2, 3 and CHECK(recover) for recid 1 (:192-199)"},
           {"msg32": arr["msg32"].hex(), "sig64": arr["sigb64"].hex(), "recover_ok": [1, 1, 1, 1],
            "note": "(r,s) = (4,4) recovers with all four recids (:176,:247-249)"}]
    return {"source": "src/modules/recovery/tests_impl.h:156-249", "vectors": vec}


def halfagg_vectors():
    """src/modules/schnorrsig_halfagg/tests_impl.h:72-167: the three half-aggregation spec vectors (all must verify)."""
    text = read("src/modules/schnorrsig_halfagg/tests_impl.h")
    body = func_body(text, "test_schnorrsig_aggverify_spec_vectors")
    arrs = parse_arrays(body)
    vec, cur = [], {}
    for name, b, _ in arrs:
        if name in cur:
            vec.append(cur); cur = {}
        cur[name] = b.hex()
        if name == "aggsig":
            vec.append(cur); cur = {}
    out = []
    for v in vec:
        out.append({"pubkeys": v.get("pubkeys_ser", ""), "msgs32": v.get("msgs32", ""), "aggsig": v["aggsig"], "expected": 1})
    assert [len(v["aggsig"]) // 64 - 1 for v in out] == [0, 1, 2]
    return {"source": "src/modules/schnorrsig_halfagg/tests_impl.h:72-167", "vectors": out}


def main():
    os.makedirs(OUT, exist_ok=True)
    from tests.refshim import RefLib
    ref = RefLib()
    jobs = {"bip340.json": bip340, "ecmult.json": ecmult_golden, "wycheproof_ecdsa.json": lambda: wycheproof(ref)}
    jobs["bppp_verify.json"] = lambda: bppp_vectors(ref)
    jobs["recovery.json"] = recovery_vectors
    jobs["halfagg.json"] = halfagg_vectors
    if hasattr(ref, "rangeproof_verify_one"):
        jobs["rangeproof.json"] = lambda: rangeproof(ref)
    if "rangeproof.json" in jobs:
        jobs["rangeproof_rewind.json"] = rewind_vectors          # after rangeproof.json (reads it)
    for name, fn in jobs.items():
        data = fn()
        with open(os.path.join(OUT, name), "w") as f:
            json.dump(data, f, indent=0, sort_keys=True)
        print(name, len(data.get("vectors", [])) or "ok")


if __name__ == "__main__":
    main()
