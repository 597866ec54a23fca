"""shardapi messages (reference src/dnet/protos/shard_api_comm.proto), see _build.py."""
from ._build import messages as _m

globals().update(_m("shard_api"))
SERVICE = "shardapi.ShardApiService"
METHODS = {n: f"/{SERVICE}/{n}" for n in ("SendFinalActivation", "SendToken")}
