"""Stand-in for ``backpack.context`` (third-party, absent from /root/reference).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference ``preds`` package import and run in
the build container so that golden vectors can be generated.  Never imported by the product.
"""


class CTX:
    active_exts = ()
    handles = []
    generation = 0
    layer_log = []        # param layers seen in the forward pass that is in flight
    pending = None        # (loss_module, logits, layers) of the last loss forward

    @staticmethod
    def remove_hooks():
        for h in CTX.handles:
            h.remove()
        CTX.handles = []
        CTX.generation += 1
        CTX.layer_log = []
        CTX.pending = None
