from .servicer import ShardApiServicer, ShardApiServer

__all__ = ["ShardApiServicer", "ShardApiServer"]
