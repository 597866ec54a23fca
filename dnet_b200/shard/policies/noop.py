from .base import ComputePolicy


class NoopPolicy(ComputePolicy):
    """Placeholder before a model is loaded (reference shard/policies/noop.py)."""

    def process(self, req):
        return None

    def configure_policy_for_model(self, req):
        return None

    def clear(self):
        return None
