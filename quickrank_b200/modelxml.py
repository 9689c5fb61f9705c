"""QuickRank XML model files <-> flat trees (Python side of the model I/O).

Schema of the reference: `Mart::save_model_to_file` (src/learning/forests/mart.cc:470-491),
`Ensemble::save_model_to_file` (src/learning/tree/ensemble.cc:133-147) and `RTNode::save_leaf`
(src/learning/tree/rtnode.cc:48-117): tab-indented, no XML declaration, 1-based `<feature>`,
thresholds printed with 9 and outputs with 17 significant digits so that float/double round-trip.
The C++ host (host/src/quickrank_host.cc) writes and parses the same files; this module lets the
ctypes binding exchange models with the reference without going through the CLI.
"""
import re

import numpy as np


def _g(v, digits):
    return "%.*g" % (digits, v)


def write_model(path, trees, weights, algo="LAMBDAMART", shrinkage=0.1, nleaves=None, nthresholds=0,
                minleafsupport=1):
    """Writes a forest of flat pre-order trees (dict(feature, threshold, left, right, value))."""
    out = ["<ranker>\n\t<info>\n"]
    nl = nleaves if nleaves is not None else max(int((t["feature"] < 0).sum()) for t in trees) if trees else 0
    out.append("\t\t<type>%s</type>\n\t\t<trees>%d</trees>\n\t\t<leaves>%d</leaves>\n" % (algo, len(trees), nl))
    out.append("\t\t<shrinkage>%s</shrinkage>\n\t\t<leafsupport>%d</leafsupport>\n" % (_g(shrinkage, 17), minleafsupport))
    out.append("\t\t<discretization>%d</discretization>\n\t\t<estop>0</estop>\n" % nthresholds)
    out.append("\t\t<subsample>1</subsample>\n\t\t<max_features>1</max_features>\n")
    out.append("\t\t<collapse_leaves_factor>0</collapse_leaves_factor>\n\t</info>\n\t<ensemble>\n")
    for i, (t, w) in enumerate(zip(trees, weights)):
        out.append("\t\t<tree id=\"%d\" weight=\"%s\">\n" % (i + 1, _g(float(w), 17)))
        feat, thr, left, right, val = t["feature"], t["threshold"], t["left"], t["right"], t["value"]
        # iterative pre-order walk; ("close", indent) entries emit the closing tags
        stack = [(0, 3, None)]
        while stack:
            node, ind, pos = stack.pop()
            tabs = "\t" * ind
            if node == "close":
                out.append(tabs + "</split>\n")
                continue
            out.append(tabs + ("<split pos=\"%s\">\n" % pos if pos else "<split>\n"))
            if feat[node] < 0:
                out.append(tabs + "\t<output>%s</output>\n" % _g(float(val[node]), 17))
                out.append(tabs + "</split>\n")
            else:
                out.append(tabs + "\t<feature>%d</feature>\n" % (int(feat[node]) + 1))
                out.append(tabs + "\t<threshold>%s</threshold>\n" % _g(float(thr[node]), 9))
                stack.append(("close", ind, None))
                stack.append((int(right[node]), ind + 1, "right"))
                stack.append((int(left[node]), ind + 1, "left"))
        out.append("\t\t</tree>\n")
    out.append("\t</ensemble>\n</ranker>\n")
    with open(path, "w") as f:
        f.write("".join(out))


_TOKEN = re.compile(r"<(/?)(\w+)([^>]*)>([^<]*)")


def read_model(path):
    """Returns (info dict, trees, weights) with trees in the flat pre-order layout."""
    text = open(path).read()
    info, trees, weights = {}, [], []
    cur = None          # arrays of the tree being read
    stack = []          # open <split> nodes: [node index, children seen]
    section = None
    for close, tag, attrs, body in _TOKEN.findall(text):
        if tag in ("info", "ensemble"):
            section = None if close else tag
        elif section == "info" and not close:
            info[tag] = body.strip()
        elif tag == "tree":
            if close:
                trees.append(dict(feature=np.array(cur["feature"], np.int32), threshold=np.array(cur["threshold"], np.float32),
                                  left=np.array(cur["left"], np.int32), right=np.array(cur["right"], np.int32),
                                  value=np.array(cur["value"], np.float64)))
                cur = None
            else:
                m = re.search(r'weight="([^"]*)"', attrs)
                weights.append(float(m.group(1)) if m else 1.0)
                cur = dict(feature=[], threshold=[], left=[], right=[], value=[])
                stack = []
        elif tag == "split" and cur is not None:
            if close:
                stack.pop()
                continue
            idx = len(cur["feature"])
            for k, v in (("feature", -1), ("threshold", 0.0), ("left", -1), ("right", -1), ("value", 0.0)):
                cur[k].append(v)
            if stack:
                parent = stack[-1]
                side = "left" if 'pos="left"' in attrs else "right" if 'pos="right"' in attrs else ("left" if parent[1] == 0 else "right")
                cur[side][parent[0]] = idx
                parent[1] += 1
            stack.append([idx, 0])
        elif cur is not None and not close and stack:
            node = stack[-1][0]
            if tag == "feature":
                cur["feature"][node] = int(body) - 1
            elif tag == "threshold":
                cur["threshold"][node] = float(np.float32(float(body)))
            elif tag == "output":
                cur["value"][node] = float(body)
    return info, trees, np.array(weights, np.float64)
