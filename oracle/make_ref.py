#!/usr/bin/env python
"""TEST INFRASTRUCTURE - generates a runnable Python-3 copy of the WISECONDOR reference into oracle/_ref/.

The reference at /root/reference is Python 2.7 (print statements, xrange, iterator .next()) and imports
packages that are absent here (pysam, matplotlib, pylab, sklearn's removed fast_dot).  This script applies
the *mechanical* source transform listed in SURVEY.md section 8(c) and writes the result to oracle/_ref/,
which is git-ignored: reference sources are never committed to this repository.  The generated tree is
used (a) here, to pin oracle/wc_oracle.py and to generate tests/golden/ fixtures, and (b) as the CPU
baseline arm of bench.py (--impl reference) when it travelled to the GPU box.

Transforms (all syntactic; no algorithmic change):
  * `print a, b`      -> `print(a, b)`; trailing comma -> end=" "; the one two-line print is joined
  * xrange -> range;  it.next() -> next(it);  fast_dot -> np.dot (sklearn removed it)
  * `map(int, ...)` argparse lambdas -> list(map(...))
Run-time shims appended to wisetools.py:
  * np.load(allow_pickle=True, encoding='latin1'); np.savez_compressed turns ragged lists into object arrays
  * PCA pinned to svd_solver='full' (what scikit-learn <= 0.17 always did; today's 'auto' picks the
    non-deterministic randomized solver for every shape used here)
  * applyPCA restated as x / ((x-mu) C^T C + mu): an unfitted sklearn PCA object no longer transforms
Stub modules: pysam, matplotlib, matplotlib.pyplot, pylab (convert/plot are out of scope).
"""
import os
import re
import sys

REF = os.environ.get("WISECONDOR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

PRINT_RE = re.compile(r"^(\s*)print (.*)$")


def _convert_prints(src):
    lines = src.split("\n")
    out = []
    i = 0
    while i < len(lines):
        line = lines[i]
        m = PRINT_RE.match(line)
        if m and not line.lstrip().startswith("#"):
            indent, body = m.group(1), m.group(2)
            # join continued statements (unbalanced brackets)
            while body.count("[") > body.count("]") or body.count("(") > body.count(")"):
                i += 1
                body += " " + lines[i].strip()
            body = body.rstrip()
            if body.endswith(","):
                out.append("%sprint(%s end=\" \")" % (indent, body))
            else:
                out.append("%sprint(%s)" % (indent, body))
        else:
            out.append(line)
        i += 1
    return "\n".join(out)


def transform(src):
    src = _convert_prints(src)
    src = src.replace("xrange(", "range(")
    src = src.replace("sam_iter.next()", "next(sam_iter)")
    src = src.replace("from sklearn.utils.extmath import fast_dot", "fast_dot = np.dot")
    src = src.replace("(lambda x: map(int,x.split(',')))", "(lambda x: list(map(int,x.split(','))))")
    return src


SHIMS = r'''

# ---- Python-3 run-time shims appended by oracle/make_ref.py (not part of the reference) ----
import functools as _functools
PCA = _functools.partial(PCA, svd_solver='full')

if not getattr(np.load, '_wc_shim', False):
	_np_load = np.load
	def _wc_load(*a, **k):
		k.setdefault('allow_pickle', True)
		k.setdefault('encoding', 'latin1')
		return _np_load(*a, **k)
	_wc_load._wc_shim = True
	np.load = _wc_load

	_np_savez = np.savez_compressed
	def _wc_savez(file, *a, **k):
		for key, val in list(k.items()):
			if isinstance(val, list) and len(val) > 0 and all(hasattr(v, 'shape') for v in val) \
					and len(set(np.shape(v) for v in val)) > 1:
				obj = np.empty(len(val), dtype=object)
				for n, v in enumerate(val):
					obj[n] = v
				k[key] = obj
		return _np_savez(file, *a, **k)
	np.savez_compressed = _wc_savez


def applyPCA(sampleData, mean, components):
	transform = np.dot(np.array([sampleData]) - mean, components.T)
	reconstructed = np.dot(transform, components) + mean
	reconstructed = reconstructed[0]
	return sampleData / reconstructed
'''

STUBS = {
    "pysam.py": "class AlignmentFile(object):\n\tdef __init__(self, *a, **k):\n\t\traise RuntimeError('pysam is not available: convert is out of scope')\n",
    "matplotlib/__init__.py": "def use(*a, **k):\n\tpass\n",
    "matplotlib/pyplot.py": "",
    "pylab.py": "def get_cmap(*a, **k):\n\treturn None\n",
}


def main():
    if not os.path.isdir(REF):
        print("reference not present at %s: nothing generated" % REF)
        return 1
    os.makedirs(os.path.join(OUT, "matplotlib"), exist_ok=True)
    for name in ("wisecondor.py", "wisetools.py", "triarray.py"):
        with open(os.path.join(REF, name)) as fh:
            src = transform(fh.read())
        if name == "wisetools.py":
            src += SHIMS
        with open(os.path.join(OUT, name), "w") as fh:
            fh.write(src)
    for name, body in STUBS.items():
        with open(os.path.join(OUT, name), "w") as fh:
            fh.write(body)
    import py_compile
    for name in ("wisecondor.py", "wisetools.py", "triarray.py"):
        py_compile.compile(os.path.join(OUT, name), doraise=True)
    print("generated", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
