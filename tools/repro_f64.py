import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rmi_b200
from tests import datasets
keys = datasets.lognormal_f64(200_000)
ds = rmi_b200.RMITrainingData(keys)
r = rmi_b200.train(ds, "radix,linear", 64)
print("ok", r.model_max_error, int(r.l1_counts.max()))
