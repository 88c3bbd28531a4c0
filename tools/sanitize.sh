#!/bin/bash
# compute-sanitizer passes over small invocations of every kernel family (race detection / sanitizers)
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from distributed_groth16_b200 import Net
from distributed_groth16_b200.dist_primitives import d_msm
from distributed_groth16_b200.groth16 import ProvingKey, prove
net = Net(0)
n = 1 << 10
b = net.generate_g1(1, n); s = net.generate_fr(2, n)
print("g1", d_msm(b, s, None, net).limbs[:2])
b2 = net.generate_g2(3, 256); s2 = net.generate_fr(4, 256)
print("g2", d_msm(b2, s2, None, net).limbs[:2])
x = net.generate_fr(5, 1 << 11).cpu().numpy().view(np.uint64)
print("ntt", net.ntt(x)[0], net.ntt(x, inverse=True, coset=True)[0])
a = net.generate_fr(6, 256).cpu().numpy().view(np.uint64)
print("h", net.h_circom(a, a, a)[0])
m, nv = 256, 200
g = lambda sd, k: net.generate_g1(sd, k).cpu().numpy().view(np.uint64)
vk1 = g(16, 3); vk2 = net.generate_g2(17, 2).cpu().numpy().view(np.uint64)
pk = ProvingKey(net, g(11, nv), g(12, nv), net.generate_g2(13, nv).cpu().numpy().view(np.uint64), g(14, nv - 2), g(15, m), 2,
                vk1[0], vk1[1], vk1[2], vk2[0], vk2[1])
z = net.generate_fr(18, nv).cpu().numpy().view(np.uint64)
print("prove", prove.create_proof(pk, z, a, a, a, z[1], z[2]).hex()[:16])
# the reference's sha256 circuit: warp-per-row QAP mat-vec, quad-cooperative bucket reduction, 16-entry tasks
from distributed_groth16_b200.groth16 import qap as qapmod
d = np.load(os.path.join(os.getcwd(), "tests", "golden", "sha256_circuit.npz"))
n_wires, n_pub, n_cons = (int(v) for v in d["dims"])
zz = net.fr_convert(net.to_device(d["witness"]), to_mont=True)
mats = qapmod.ConstraintMatrices(net, n_pub + 1, n_cons, (d["a_rows"], d["a_cols"], d["a_vals"]),
                                 (d["b_rows"], d["b_cols"], d["b_vals"]), values_montgomery_depth=-1)
q = qapmod.qap(mats, zz, net)
print("qap", q.a[0].cpu().numpy()[:1])
bw = net.generate_g2(21, int(zz.shape[0]))
print("g2 over the sha256 witness", d_msm(bw, zz, None, net).limbs[:2])
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san_small.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "$tool exit=$?" >> gpurun_out/sanitizer_$tool.txt
  tail -4 gpurun_out/sanitizer_$tool.txt
done
