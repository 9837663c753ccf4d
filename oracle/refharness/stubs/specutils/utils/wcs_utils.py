def vac_to_air(*a, **k):
    raise NotImplementedError('stub specutils')


def air_to_vac(*a, **k):
    raise NotImplementedError('stub specutils')
