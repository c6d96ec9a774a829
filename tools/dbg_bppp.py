import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import secp256k1_zkp_b200 as b200
d = json.load(open("tests/golden/bppp_verify.json"))
gens = bytes.fromhex(d["gens_xy"])
fb = np.load("tests/golden/bppp_fixture.npz")
for gw in (16, 12, 8):
    eng = b200.Engine(gen_window_bits=gw)
    res = []
    for v in d["vectors"]:
        g, h = v["g_len"], v["h_len"]
        if g == 0: continue
        pr = np.frombuffer(bytes.fromhex(v["proof"]), dtype=np.uint8).reshape(1, -1)
        cm, rho, cv = bytes.fromhex(v["commit_xy"]), bytes.fromhex(v["r32"]), bytes.fromhex(v["c_vec32"])
        tr = eng.bppp_transcripts(cm, rho, gens[:64 * (g + h)], g, cv, h, mode="plain")
        got = eng.bppp_norm_verify(pr, [pr.shape[1]], tr, rho, gens[:64 * (g + h)], g, cv, h, cm)
        res.append((v["index"], int(got[0]), v["expected"]))
    print("gw", gw, "golden mismatches:", [r for r in res if r[1] != r[2]])
    tr = eng.bppp_transcripts(fb["commits"], fb["rho"], fb["gens"], 16, fb["c_vec"], 8, mode="tests")
    for rep in range(2):
        ok = eng.bppp_norm_verify(fb["proofs"], [324] * 64, tr, fb["rho"], fb["gens"], 16, fb["c_vec"], 8, fb["commits"])
        print("  fixture accepted", int(ok.sum()), "of 64")
    eng.close()
