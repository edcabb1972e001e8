import os, sys, time
sys.path.insert(0, '/root/repo')
import torch, lambdaworks_kzg_b200 as lw
n = 1024
lw.set_option("window_bits", 10)
s = lw.load_trusted_setup_file('/root/repo/tests/golden/trusted_setup.txt')
blobs = b"".join(lw.synth_blob_host(k) for k in range(n))
coms, proofs, st = lw.commit_and_prove_batch(blobs, n, s)
B = 131072
bl = [blobs[i*B:(i+1)*B] for i in range(n)]
print(lw.verify_blob_kzg_proof_batch(bl, coms, proofs, s))
print(lw.verify_kzg_proof(coms[0], bytes(32), bl[0][:32], bytes([0xC0]) + bytes(47), s))
