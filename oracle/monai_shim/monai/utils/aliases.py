def alias(*names):
    def deco(obj):
        return obj
    return deco
