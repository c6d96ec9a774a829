import os, subprocess, sys
for name in sorted(os.listdir("abtest")):
    env = dict(os.environ, SECP256K1_B200_LIB=os.path.abspath(os.path.join("abtest", name)))
    out = subprocess.run([sys.executable, "tools/dbg_bppp2.py"], env=env, capture_output=True, text=True).stdout.split("\n")
    print(name, [l for l in out if l][::3])
