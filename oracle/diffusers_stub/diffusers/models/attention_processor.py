class SpatialNorm:      # only instantiated for norm_type == "spatial", which no iVideoGPT config uses
    def __init__(self, *a, **k):
        raise NotImplementedError("SpatialNorm is not part of the iVideoGPT configs (norm_type 'group')")
