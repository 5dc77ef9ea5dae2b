"""Print where evaluate_device and the reference fixture disagree (debug aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rsgnet_b200 import synth
from rsgnet_b200.nms.nms import evaluate_device
from oracle import nms_oracle
g = np.load(os.path.join(ROOT, 'tests/golden/evaluate_crowdpose.npz'))
preds, boxes, ids = synth.evaluate_inputs(400, 12, 14, seed=13)
out = evaluate_device(preds, boxes, ids, 0.9, 0.2, nms_oracle.CROWDPOSE_SIGMAS).host()
off = np.concatenate([[0], np.cumsum(g['counts'])])
nbad = 0
sc = np.array([nms_oracle.rescore(boxes[i, 5], preds[i, :, 2], 0.2) for i in range(len(ids))])
print('scores equal', np.array_equal(sc, out['scores']), np.abs(sc - out['scores']).max())
for r in range(len(g['images'])):
    a, b = out['keep'][off[r]:off[r + 1]], g['keep'][off[r]:off[r + 1]]
    if not np.array_equal(a, b):
        nbad += 1
        if nbad <= 5:
            m = np.nonzero(ids == g['images'][r])[0]
            print('image', r, 'n', len(m), 'got', list(a), 'ref', list(b), 'scores', [float(sc[i]) for i in m], 'members', list(m))
print('bad images', nbad)
