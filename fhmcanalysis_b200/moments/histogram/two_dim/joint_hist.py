"""Drop-in replacement for FHMCAnalysis.moments.histogram.two_dim.joint_hist (reference file
moments/histogram/two_dim/joint_hist.pyx, "JH"): container of a joint two-dimensional histogram lnPI(op1, op2)
assembled from one-dimensional entries, plus the NEW batched reweighting of that surface on the GPU (K5).

The container part (entry / add / enter / make / to_json / from_json) is bookkeeping of small ragged lists and
keeps the reference's layout exactly: ``data['ln(PI)']`` dense and padded with -inf, ``data['bounds_idx'][j] =
[first, last]`` INCLUSIVE column indices of row j, ``data['props'][name]`` dense and padded with 0 (JH:201-247).
"""
import copy
import json

import numpy as np

from fhmcanalysis_b200 import engine


class joint_hist(object):
    """Joint histogram (JH:22-301)."""

    class entry(object):
        """One lnPI(op2) vector at fixed op1 (JH:28-138)."""

        def __init__(self):
            self.clear_all()

        def clear_all(self):
            self.data = {}

        def clear_props(self):
            self.data["props"] = {}

        def set(self, lnpi, op_vals, name_val_dict):
            self.set_lnpi(lnpi, op_vals)
            for p in name_val_dict:
                self.set_prop(p, name_val_dict[p])

        def set_lnpi(self, lnpi, op_vals):
            assert len(op_vals) == len(lnpi), "Size mismatch between ln(PI) and order parameters"
            self.data["ln(PI)"] = np.array(lnpi, dtype=np.float64)
            assert np.all(sorted(op_vals) == op_vals), "Order parameter values are not sorted"
            self.data["op_vals"] = np.array(op_vals, dtype=np.float64)
            if "props" in self.data:
                for x in self.data["props"]:
                    assert self._check_size(self.data["props"][x]), \
                        "Size of existing properties vectors is different from new ln(PI)"

        def set_prop(self, name, val):
            if "props" not in self.data:
                self.data["props"] = {}
            assert self._check_size(val), "Size of property is different from ln(PI)"
            self.data["props"][name] = np.array(val, dtype=np.float64)

        def _check_size(self, x):
            ref_size = len(self.data["ln(PI)"]) if "ln(PI)" in self.data else len(x)
            return len(x) == ref_size

    def __init__(self):
        self.clear()

    def clear(self):
        self.data = {}

    def add(self, op1, entry):
        if "entries" not in self.data:
            self.data["entries"] = {}
        self.data["entries"][op1] = copy.deepcopy(entry)

    def enter(self, op1, lnpi, op_vals, name_val_dict):
        e = self.entry()
        e.set(lnpi, op_vals, name_val_dict)
        self.add(op1, e)

    def make(self):
        """Sort all raw entries into a self-consistent dense surface (JH:201-247)."""
        op1_vals = sorted(self.data["entries"])
        op2_set = set()
        for x in op1_vals:
            op2_set |= set(self.data["entries"][x].data["op_vals"])
        op2_vals = sorted(op2_set)
        col = {v: i for i, v in enumerate(op2_vals)}
        n1, n2 = len(op1_vals), len(op2_vals)
        self.data["ln(PI)"] = np.full((n1, n2), -np.inf, dtype=np.float64)
        self.data["op_1"] = np.array(op1_vals, dtype=np.float64)
        self.data["op_2"] = np.array(op2_vals, dtype=np.float64)
        self.data["bounds_idx"] = np.full((n1, 2), 0, dtype=np.int64)
        self.data["props"] = {}
        all_props = []
        for j, x in enumerate(op1_vals):
            e = self.data["entries"][x].data
            idx = np.array([col[v] for v in e["op_vals"]], dtype=np.int64)
            self.data["ln(PI)"][j, idx] = e["ln(PI)"]
            self.data["bounds_idx"][j, :] = [idx.min(), idx.max()]
            props = sorted(e.get("props", {}))
            if len(all_props) > 0:
                assert props == all_props, "Properties are not all the same, or some are missing"
            else:
                all_props = copy.copy(props)
        for prop in all_props:
            self.data["props"][prop] = np.full((n1, n2), 0, dtype=np.float64)
            for j, x in enumerate(op1_vals):
                e = self.data["entries"][x].data
                idx = np.array([col[v] for v in e["op_vals"]], dtype=np.int64)
                self.data["props"][prop][j, idx] = e["props"][prop]

    def to_json(self, fname):
        obj = copy.deepcopy(self.data)
        obj.pop("entries", None)
        for k in ("ln(PI)", "op_1", "op_2", "bounds_idx"):
            obj[k] = obj[k].tolist()
        for p in obj["props"]:
            obj["props"][p] = obj["props"][p].tolist()
        with open(fname, "w") as f:
            json.dump(obj, f, indent=4, sort_keys=True)

    def from_json(self, fname):
        self.clear()
        with open(fname, "r") as f:
            raw = json.load(f)
        for k, msg in (("ln(PI)", "ln(PI)"), ("op_1", "op_1"), ("op_2", "op_2"), ("bounds_idx", "bounds"), ("props", "properties")):
            assert k in raw, "Missing " + msg + " information"
        self.data["ln(PI)"] = np.array(raw["ln(PI)"], dtype=np.float64)
        self.data["op_1"] = np.array(raw["op_1"], dtype=np.float64)
        self.data["op_2"] = np.array(raw["op_2"], dtype=np.float64)
        self.data["bounds_idx"] = np.array(raw["bounds_idx"], dtype=np.float64)
        self.data["props"] = {p: np.array(raw["props"][p], dtype=np.float64) for p in raw["props"]}

    # ------------------------------------------------------------------------------------------
    # new: batched reweighting of the surface (K5)
    # ------------------------------------------------------------------------------------------
    def reweight_batch(self, a1, a2, props=(), device=None, return_device=False):
        """Reweight lnPI(op1, op2) to many state points at once:

            v_ij(s) = lnPI_ij + a1[s]*op1_i + a2[s]*op2_j     over the support of every row,

        e.g. (op1, op2) = (N_1, N_2) and a_k = beta*(mu_k - mu_k,ref).  Returns an array [S, 3 + len(props)]:
        ln sum exp v, <op1>, <op2>, <prop> ... (at most two properties per call)."""
        if "ln(PI)" not in self.data:
            raise Exception("call make() (or from_json()) before reweighting")
        b = np.asarray(self.data["bounds_idx"]).astype(np.int32).copy()
        b[:, 1] += 1  # reference bounds are inclusive (JH:232-238); the kernel takes [lo, hi)
        pr = None
        if len(props):
            pr = np.stack([np.asarray(self.data["props"][p], dtype=np.float64) for p in props])
        return engine.reweight_2d(self.data["ln(PI)"], b, self.data["op_1"], self.data["op_2"], a1, a2, pr,
                                  device=device, return_device=return_device)


if __name__ == "__main__":
    print("joint_hist (B200)")
