"""Run the reference's ROS node (scripts/taichislam_node.py) - or any other script of a TaichiSLAM checkout -
UNMODIFIED on this backend:

    python -m taichislam_b200.run_node /path/to/TaichiSLAM/scripts/taichislam_node.py [node args ...]

The node script puts its own checkout first on sys.path (taichislam_node.py:4), so a PYTHONPATH entry alone cannot
redirect `taichi_slam.mapping`.  This launcher therefore imports the alias package `taichi_slam` of this repo FIRST -
the import system then finds `taichi_slam.mapping[.dense_tsdf ...]` in it (-> taichislam_b200.mapping) whatever
sys.path says later - with the checkout's own `taichi_slam` directory appended to the package's `__path__`
(pkgutil.extend_path), so everything the backend does not replace (`taichi_slam.utils.*`: ROS / rendering / LCM glue,
`taichi_slam.taichi_opti`) still comes from the checkout.  Then the script runs as `__main__`.
"""
import os
import runpy
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def prepare(checkout_root):
    """Make `taichi_slam` resolve to the alias package, backed by `checkout_root` for the modules it does not provide."""
    checkout_root = os.path.abspath(checkout_root)
    for k in [k for k in sys.modules if k == "taichi_slam" or k.startswith("taichi_slam.")]:
        del sys.modules[k]
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    if checkout_root not in sys.path:
        sys.path.append(checkout_root)  # must be visible while the alias package extends its __path__
    import taichi_slam
    if not any(os.path.abspath(p).startswith(_REPO) for p in taichi_slam.__path__):
        raise ImportError(f"taichi_slam resolved to {taichi_slam.__path__}, not to the alias package of {_REPO}")
    return taichi_slam


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        sys.stderr.write(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    prepare(os.path.dirname(os.path.dirname(script)))  # <checkout>/scripts/<node>.py
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
