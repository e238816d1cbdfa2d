def apply_forward_hook(method):       # a no-op unless an accelerate offload hook is attached
    return method
