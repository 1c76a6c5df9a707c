import numpy as np, sys
sys.path.insert(0,'.')
import gisnav_b200
from gisnav_b200 import synth, weights as W
blob=W.pack(W.random_init(0))
ctx=gisnav_b200.Context(gisnav_b200.Config(max_batch=1,max_image_h=256,max_image_w=320,max_keypoints=64), weights=blob)
img=np.ascontiguousarray(synth.ground_texture(512, seed=11, n_shapes=300)[40:136, 60:188])
xy,sc,d=gisnav_b200.KeypointExtractor(ctx).detect_and_compute_arrays(img)
print(len(xy))
