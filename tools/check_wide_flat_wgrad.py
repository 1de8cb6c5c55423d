import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import selftest_conv as st
ok = True
ok &= st.case_wgrad("k3s1", 64, 16, 128, 128)
ok &= st.case_wgrad("k3s1", 32, 16, 256, 128)
ok &= st.case_wgrad("k3s1", 16, 32, 64, 64)
ok &= st.case_wgrad("k3s1", 50, 8, 192, 192)
print("WIDE_FLAT", "OK" if ok else "FAILED")
