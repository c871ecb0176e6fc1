"""Marker classes standing in for ``backpack.extensions`` (see context.py header)."""


class _Ext:
    second_order = False


class BatchGrad(_Ext):
    pass


class DiagGGNExact(_Ext):
    second_order = True


class KFLR(_Ext):
    second_order = True


class KFAC(_Ext):
    second_order = True
