import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'sudoku-vision_b200'))
import numpy as np, torch
from svb200 import Scanner
from oracle import oracle as O
sc = Scanner()
for hw, n in [((72,480),1),((300,1920),2)]:
    rng = np.random.default_rng(hw[0]+hw[1])
    img = rng.integers(0,256,(n,)+hw+(3,)).astype(np.uint8)
    m = sc.preprocess(torch.from_numpy(img).cuda()).cpu().numpy()
    for i in range(n):
        want = O.preprocess(img[i]); bad = np.argwhere(m[i]!=want)
        rows, rc = np.unique(bad[:,0], return_counts=True) if len(bad) else ([],[])
        print(hw, i, 'nbad', len(bad), 'of', want.size, 'rows', list(zip(rows[:20].tolist(), rc[:20].tolist())), 'cols%4 hist', np.bincount(bad[:,1]%4, minlength=4).tolist() if len(bad) else None)
        if len(bad):
            y,x = bad[0]; print('  first', y, x, 'got', m[i][y,x], 'want', want[y,x], 'gray/blur around', O.blur5(O.gray(img[i]))[y, max(x-2,0):x+3].tolist())
