import importlib, time, numpy as np, os, sys
sys.path.insert(0, os.getcwd())
lib = importlib.import_module("6dpose_b200._lib")
gold = np.load("tests/golden/icp_case1.npz")
icp = lib.NativeIcp(0)
for nh in (1, 3, 8):
    xy = [[int(v) for v in gold["xy_shift_0"]]] * nh
    a = dict(scene_depth=gold["scene_shift_0"], model_depths=[gold["model"]] * nh, sceneK=gold["K"],
             modelKs=np.stack([gold["K"]] * nh), Rs=np.stack([gold["R"]] * nh), ts=np.stack([gold["t"].reshape(3)] * nh), detect_xy=xy)
    for _ in range(3): icp.process_batch(**a)
    t0 = time.perf_counter()
    for _ in range(20): Ro, to, res = icp.process_batch(**a)
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(nh, "hyp: %.3f ms/call" % ms, icp.last_stats(), res[:2])
