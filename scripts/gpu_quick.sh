cd /root/repo
python - <<'PY'
import numpy as np, sys
sys.path.insert(0,'.')
from jax_sph_b200 import Engine, config_from_setup
from oracle import cases
s3 = cases.make_case("tgv", dim=3, dx=2*np.pi/32, dtype=np.float32, tvf=1.0, viscosity=0.02)
eng = Engine(config_from_setup(s3), len(s3.state["r"]))
eng.upload(s3.state)
try:
    eng.step(s3.dt, 2)
    print("ok", eng.error())
except Exception as e:
    print("ERR", e)
PY
