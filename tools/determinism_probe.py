import sys, importlib, numpy as np, torch, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
emd = importlib.import_module("ai-cv-automation-elect-micr_b200")
rng = np.random.default_rng(99)
crops = rng.random((16, 512, 512)).astype(np.float32)
eng = emd.Engine(cropsize=512, max_batch=16)
eng.load_weights(emd.weights.pack(emd.weights.init_reference_weights(1)))
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
a = eng.forward(crops, mode=mode)                                   # host path: 2 chunks of 8
d = eng.forward(torch.from_numpy(crops).cuda(), mode=mode); torch.cuda.synchronize(); d = d.cpu().numpy()   # one pass of 16
d2 = eng.forward(torch.from_numpy(crops).cuda(), mode=mode); torch.cuda.synchronize(); d2 = d2.cpu().numpy()
print(mode, "host-vs-dev mismatches per crop:", [(int((a[i] != d[i]).sum())) for i in range(16)])
print(mode, "dev-vs-dev  mismatches per crop:", [(int((d2[i] != d[i]).sum())) for i in range(16)])
# which layer first differs between an 8-batch and a 16-batch pass?
eng.set_keep_activations(True)
x16 = torch.from_numpy(crops).cuda()
eng.forward(x16, mode=mode); torch.cuda.synchronize()
names = ["cnn0","cnn0_last","enc0","cnn1","enc1","enc2","enc3","trunk4","trunk_mid0","trunk_mid5","trunk_mid10","aspp_1x1","aspp_r6","aspp_pellet","upsample4","deconv2_0","dec2","deconv2to1","deconv1_0","dec1","deconv1to0","deconv0_0","residual0_d","dec0"]
A = {n: eng.activation(n).copy() for n in names}
x8 = torch.from_numpy(crops[:8]).cuda()
eng.forward(x8, mode=mode); torch.cuda.synchronize()
for n in names:
    b = eng.activation(n)
    print(f"{n:14s} mismatches {(b != A[n][:8]).sum()} of {b.size}")
