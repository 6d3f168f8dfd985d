# Builds the CPU oracle's C port (TEST INFRASTRUCTURE) into oracle/_build/libmsmref.so.
# There is no oracle/_ref: the reference is Go and no Go toolchain exists in this image.
CC ?= gcc
CFLAGS ?= -O3 -march=native -fPIC -std=gnu11 -Wall -Wno-unused-function
all: _build/libmsmref.so
consts.h: gen_consts.py
	python gen_consts.py
_build/libmsmref.so: msm_ref.c fp_tmpl.h fp2_tmpl.h group_tmpl.h consts.h
	mkdir -p _build
	$(CC) $(CFLAGS) -shared -o $@ msm_ref.c -lpthread
clean:
	rm -rf _build
