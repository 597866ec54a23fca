from .servicer import GrpcServicer
from .server import GrpcServer

__all__ = ["GrpcServicer", "GrpcServer"]
