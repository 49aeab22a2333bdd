class ProfileOptions:
    pass
