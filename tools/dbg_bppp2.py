import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import secp256k1_zkp_b200 as b200
eng = b200.Engine(gen_window_bits=4)
d = json.load(open("tests/golden/bppp_verify.json"))
gens = bytes.fromhex(d["gens_xy"])
for idx in (0, 6, 9):
    v = d["vectors"][idx]
    g, h = v["g_len"], v["h_len"]
    pr = np.frombuffer(bytes.fromhex(v["proof"]), dtype=np.uint8).reshape(1, -1)
    cm, rho, cv = bytes.fromhex(v["commit_xy"]), bytes.fromhex(v["r32"]), bytes.fromhex(v["c_vec32"])
    tr = eng.bppp_transcripts(cm, rho, gens[:64 * (g + h)], g, cv, h, mode="plain")
    for rep in range(3):
        got = eng.bppp_norm_verify(pr, [pr.shape[1]], tr, rho, gens[:64 * (g + h)], g, cv, h, cm)
        print(idx, "got", int(got[0]), "exp", v["expected"])
