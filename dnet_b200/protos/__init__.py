"""Wire contract (reference src/dnet/protos/*.proto), built without protoc."""
