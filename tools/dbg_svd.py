import sys
import numpy as np
sys.path.insert(0, ".")
from itensorsgpu_b200 import tn
rng = np.random.default_rng(41)
for shape in [(130, 57), (113, 113), (200, 113), (300, 170), (300, 260), (301, 260), (512, 512)]:
    A = rng.standard_normal(shape)
    U, S, V, err = tn.ops.svd(tn.DTensor.from_numpy(A))
    U, S, V = U.numpy(), S.cpu().numpy(), V.numpy()
    Sr = np.linalg.svd(A, compute_uv=False)
    print(shape, "rec", np.linalg.norm(U @ np.diag(S) @ V.T - A) / np.linalg.norm(A), "U", np.linalg.norm(U.T @ U - np.eye(len(S))),
          "V", np.linalg.norm(V.T @ V - np.eye(len(S))), "S", np.max(np.abs(S - Sr)), flush=True)
