/* stand-in: everything is declared in the stub GL/gl.h */
