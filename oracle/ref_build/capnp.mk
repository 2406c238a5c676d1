# Builds the Cap'n Proto runtime + schema compiler that the reference vendors
# (/root/reference/src/3rdparty/capnproto) straight from its sources with g++ --
# no cmake, nothing copied into this repo.  Outputs go to oracle/_ref/capnp/.
# TEST INFRASTRUCTURE ONLY (see oracle/README.md).
REF   ?= /root/reference
CSRC  := $(REF)/src/3rdparty/capnproto/c++/src
OUT   ?= $(abspath $(dir $(lastword $(MAKEFILE_LIST)))/../_ref)/capnp
CXX   ?= g++
CXXFLAGS := -std=c++20 -O2 -fPIC -w -I$(CSRC) -DCAPNP_INCLUDE_DIR=\"$(CSRC)\" -DVERSION=\"vendored\"

KJ := array cidr list common debug exception io memory mutex string source-location hash table thread \
      main arena units encoding refcount string-tree time filesystem filesystem-disk-unix parse/char
CAPNP := c++.capnp blob arena layout list any message schema.capnp stream.capnp serialize serialize-packed \
      schema schema-loader dynamic stringify
CAPNPC := compiler/type-id compiler/error-reporter compiler/lexer.capnp compiler/lexer compiler/grammar.capnp \
      compiler/parser compiler/generics compiler/node-translator compiler/compiler schema-parser serialize-text
JSON := compat/json compat/json.capnp

KJ_O     := $(addprefix $(OUT)/obj/kj/,$(addsuffix .o,$(KJ)))
CAPNP_O  := $(addprefix $(OUT)/obj/capnp/,$(addsuffix .o,$(CAPNP)))
CAPNPC_O := $(addprefix $(OUT)/obj/capnp/,$(addsuffix .o,$(CAPNPC) $(JSON)))

all: $(OUT)/libcapnp_kj.a $(OUT)/bin/capnp $(OUT)/bin/capnpc-c++

$(OUT)/obj/kj/%.o: $(CSRC)/kj/%.c++
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OUT)/obj/capnp/%.o: $(CSRC)/capnp/%.c++
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(OUT)/libcapnp_kj.a: $(KJ_O) $(CAPNP_O)
	ar rcs $@ $^
$(OUT)/libcapnpc.a: $(CAPNPC_O)
	ar rcs $@ $^
$(OUT)/bin/capnp: $(OUT)/obj/capnp/compiler/module-loader.o $(OUT)/obj/capnp/compiler/capnp.o $(OUT)/libcapnpc.a $(OUT)/libcapnp_kj.a
	@mkdir -p $(dir $@)
	$(CXX) -o $@ $(OUT)/obj/capnp/compiler/module-loader.o $(OUT)/obj/capnp/compiler/capnp.o $(OUT)/libcapnpc.a $(OUT)/libcapnp_kj.a -lpthread
$(OUT)/bin/capnpc-c++: $(OUT)/obj/capnp/compiler/capnpc-c++.o $(OUT)/libcapnp_kj.a
	@mkdir -p $(dir $@)
	$(CXX) -o $@ $< $(OUT)/libcapnp_kj.a -lpthread
.PHONY: all
