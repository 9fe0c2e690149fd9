"""Importable alias: the package directory name required by the build contract contains hyphens
(`adapting-rgb-pose-estimation-to-new-domains_b200`), which `import` cannot spell; this module
loads it with importlib and registers it as `rmpe_b200` (and its submodules as
`rmpe_b200.<sub>`), so `import rmpe_b200.batch` etc. work."""
import importlib
import os
import sys

_PKG = "adapting-rgb-pose-estimation-to-new-domains_b200"
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module(_PKG)


def sub(name):
    """rmpe_b200.sub('batch') -> the package's submodule."""
    return importlib.import_module(_PKG + "." + name)


batch = sub("batch")
synth = sub("synth")
lib = sub("_lib")
config = sub("py_rmpe_server.py_rmpe_config")
transformer = sub("py_rmpe_server.py_rmpe_transformer")
heatmapper = sub("py_rmpe_server.py_rmpe_heatmapper")
data_iterator = sub("py_rmpe_server.py_rmpe_data_iterator")
decode = sub("eval.eval_coco2014_multi_modes")
util = sub("util")
ds_generators = sub("training.ds_generators")
package = _pkg
