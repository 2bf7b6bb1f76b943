"""Template bank: storage, OpenCV-FileStorage YAML IO, and flat packing for the C-ABI.

Mirrors (reference: linemodLevelup/linemodLevelup.{h,cpp}, "LL.h"/"LL.cpp"):
  Feature / Template / TemplatePyramid / TemplatesMap     LL.h:23-45, 361-362
  Template::read/write, Feature::read/write               LL.cpp:194-232
  Detector::readClass / writeClass                        LL.cpp:2043-2122

A template pyramid is a list of L*M templates ordered [level*M + modality] (LL.cpp:1964).
Each template is (width, height, pyramid_level, features int32 [n,3] = x,y,label).
"""
import re

import numpy as np

MODALITY_NAMES = ["ColorGradient", "DepthNormal"]


class Template:
    __slots__ = ("width", "height", "pyramid_level", "features")

    def __init__(self, width=0, height=0, pyramid_level=0, features=None):
        self.width = int(width)
        self.height = int(height)
        self.pyramid_level = int(pyramid_level)
        self.features = np.zeros((0, 3), np.int32) if features is None else np.ascontiguousarray(features, np.int32).reshape(-1, 3)


class TemplateBank:
    """class_id -> list of template pyramids.  Iteration order for matching = std::map order
    (sorted class ids), as in LL.cpp:1756-1758."""

    def __init__(self):
        self.classes = {}

    def num_templates(self, class_id=None):
        if class_id is not None:
            return len(self.classes.get(class_id, []))
        return sum(len(v) for v in self.classes.values())

    def class_ids(self):
        return sorted(self.classes.keys())

    # ---- YAML (OpenCV FileStorage dialect, YAML 1.0) -------------------------------------
    _tok = re.compile(
        r"^\s*(?:-\s*)?(?:"
        r"(?P<key>class_id|pyramid_levels|template_id|width|height|pyramid_level|modalities)\s*:\s*(?P<val>.*?)\s*$"
        r"|\[\s*(?P<x>-?\d+)\s*,\s*(?P<y>-?\d+)\s*,\s*(?P<l>-?\d+)\s*\]\s*$)",
        re.M)

    def read_class(self, path, expected_levels, class_id_override=""):
        """Detector::readClass, LL.cpp:2043-2091.  Raises RuntimeError where the reference CV_Asserts."""
        import gzip
        opener = gzip.open if path.endswith(".gz") else open
        try:
            with opener(path, "rt") as fh:
                text = fh.read()
        except OSError as e:
            raise RuntimeError("cannot open template file %s: %s" % (path, e))
        class_id = None
        levels = None
        pyramids = []
        cur_tp = None
        cur_t = None
        feats = None
        for m in self._tok.finditer(text):
            if m.group("x") is not None:
                feats.append((int(m.group("x")), int(m.group("y")), int(m.group("l"))))
                continue
            key, val = m.group("key"), m.group("val")
            if key == "class_id":
                class_id = val.strip().strip('"')
            elif key == "modalities":
                names = [s.strip() for s in val.strip("[] ").split(",") if s.strip()]
                if names != MODALITY_NAMES:  # LL.cpp:2047-2051
                    raise RuntimeError("modalities mismatch: %r" % (names,))
            elif key == "pyramid_levels":
                levels = int(val)
            elif key == "template_id":
                if int(val) != len(pyramids):  # LL.cpp:2077
                    raise RuntimeError("template_id == expected_id")
                cur_tp = []
                pyramids.append(cur_tp)
            elif key == "width":
                cur_t = Template(width=int(val))
                feats = []
                cur_t.features = feats
                cur_tp.append(cur_t)
            elif key == "height":
                cur_t.height = int(val)
            elif key == "pyramid_level":
                cur_t.pyramid_level = int(val)
        if levels != expected_levels:  # LL.cpp:2052
            raise RuntimeError("pyramid_levels mismatch: file %r detector %r" % (levels, expected_levels))
        if class_id_override:
            class_id = class_id_override
        elif class_id in self.classes:  # LL.cpp:2059
            raise RuntimeError("class %s already loaded" % class_id)
        for tp in pyramids:
            for t in tp:
                t.features = np.asarray(t.features, np.int32).reshape(-1, 3)
        self.classes[class_id] = pyramids
        return class_id

    def write_class(self, class_id, path, levels):
        """Detector::writeClass, LL.cpp:2093-2122 (same node layout as the reference's files)."""
        import gzip
        tps = self.classes[class_id]
        out = ["%YAML:1.0", "---", 'class_id: "%s"' % class_id,
               "modalities: [ %s ]" % ", ".join(MODALITY_NAMES), "pyramid_levels: %d" % levels,
               "template_pyramids:"]
        for i, tp in enumerate(tps):
            out.append("   -")
            out.append("      template_id: %d" % i)
            out.append("      templates:")
            for t in tp:
                out.append("         -")
                out.append("            width: %d" % t.width)
                out.append("            height: %d" % t.height)
                out.append("            pyramid_level: %d" % t.pyramid_level)
                out.append("            features:")
                for x, y, l in t.features.tolist():
                    out.append("               - [ %d, %d, %d ]" % (x, y, l))
        data = "\n".join(out) + "\n"
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "wt") as fh:
            fh.write(data)

    # ---- flat packing ------------------------------------------------------------------
    def pack(self, class_ids, slots):
        """Flatten the given classes (in the given order) for the C-ABI / oracle.

        Returns dict(class_begin int32 [C+1], tmeta int32 [G, slots, 4] = width,height,feat_begin,feat_count,
        feats int32 [F,3]).  slots = pyramid_levels * modalities.
        """
        class_begin = [0]
        metas = []
        feats = []
        nfe = 0
        for cid in class_ids:
            tps = self.classes[cid]
            for tp in tps:
                if len(tp) != slots:
                    raise RuntimeError("template pyramid of class %s has %d templates, expected %d" % (cid, len(tp), slots))
                row = []
                for t in tp:
                    n = int(t.features.shape[0])
                    row.append((t.width, t.height, nfe, n))
                    feats.append(t.features)
                    nfe += n
                metas.append(row)
            class_begin.append(len(metas))
        tmeta = np.asarray(metas, np.int32).reshape(len(metas), slots, 4)
        allf = np.concatenate(feats, 0).astype(np.int32) if feats else np.zeros((0, 3), np.int32)
        return dict(class_begin=np.asarray(class_begin, np.int32), tmeta=np.ascontiguousarray(tmeta),
                    feats=np.ascontiguousarray(allf))
