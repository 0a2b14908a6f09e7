import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'sudoku-vision_b200'))
import numpy as np, torch
from svb200 import Scanner
from oracle import oracle as O
sc = Scanner()
for hw, n in [((64,64),4),((72,480),3),((300,1920),3)]:
    rng = np.random.default_rng(hw[0]+hw[1])
    img = rng.integers(0,256,(n,)+hw+(3,)).astype(np.uint8)
    out = torch.full((n,)+hw, 77, dtype=torch.uint8, device='cuda')
    m = sc.preprocess(torch.from_numpy(img).cuda(), out=out).cpu().numpy()
    for i in range(n):
        want = O.preprocess(img[i]); bad = np.argwhere(m[i]!=want)
        vals = np.unique(m[i][m[i]!=want]) if len(bad) else []
        print(hw, i, 'nbad', len(bad), 'rows', np.unique(bad[:,0])[:12].tolist(), 'bad values', list(vals)[:5], 'count77', int((m[i]==77).sum()))
