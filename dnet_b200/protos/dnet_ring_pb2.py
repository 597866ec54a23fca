"""dnetring messages (reference src/dnet/protos/dnet_ring.proto), see _build.py."""
from ._build import messages as _m

globals().update(_m("ring"))
SERVICE = "dnetring.DnetRingService"
METHODS = {n: f"/{SERVICE}/{n}" for n in ("SendActivation", "HealthCheck", "ResetCache", "MeasureLatency", "StreamActivations")}
