// stand-in: the viewer header only needs the include to exist
