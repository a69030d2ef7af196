class UNet2DConditionLoadersMixin:
    pass


class FromOriginalControlnetMixin:
    pass
