# Builds the product library (CUDA, sm_100a only) and the test-only oracle libraries.
#   make            -> eao-fusion_b200/lib/libeaof_orb.so
#   make oracle     -> oracle/liborb_oracle.so (+ oracle/_ref/liborb_ref.so when /root/reference is present)
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC,-Wall -Xptxas -v $(EXTRA)
PKG       := eao-fusion_b200
LIB       := $(PKG)/lib/libeaof_orb.so
SRCS      := $(PKG)/csrc/eaof_orb.cu $(PKG)/csrc/eaof_match.cu $(PKG)/csrc/eaof_voc.cu $(PKG)/csrc/eaof_sweep.cu
HDRS      := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.h) include/eaof_orb.h include/eaof_match.h include/eaof_voc.h include/eaof_sweep.h

DROPIN_T  := tests/cpp/_build/libdropin_harness.so
DROPIN_M  := tests/cpp/_build/libmatch_dropin.so
DROPIN_V  := tests/cpp/_build/libvoc_dropin.so
REF       ?= /root/reference

all: $(LIB) $(DROPIN_T) matchdropin vocdropin

$(LIB): $(SRCS) $(HDRS)
	@mkdir -p $(PKG)/lib build
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(SRCS) -ldl 2> build/ptxas.log || (cat build/ptxas.log; exit 1)
	@grep -E "registers|spill|error" build/ptxas.log | sed 's/^/  /' | head -80

# the drop-in ORB_SLAM2::ORBextractor compiled against the cv shim (test harness; deployment compiles
# $(PKG)/dropin/ORBextractor.cc against the real OpenCV)
$(DROPIN_T): tests/cpp/dropin_harness.cc $(PKG)/dropin/ORBextractor.cc $(PKG)/dropin/ORBextractor.h include/eaof_orb.h $(LIB)
	@mkdir -p tests/cpp/_build
	g++ -O2 -std=c++14 -fPIC -shared -Wall -Wl,-Bsymbolic -Ioracle/cvshim -I$(PKG)/dropin -Iinclude \
	    -o $@ tests/cpp/dropin_harness.cc $(PKG)/dropin/ORBextractor.cc -L$(PKG)/lib -leaof_orb -Wl,-rpath,'$$ORIGIN/../../../$(PKG)/lib'

# the drop-in ORB_SLAM2::ORBmatcher methods ($(PKG)/dropin/ORBmatcher.cc, compiled against the reference's own
# include/ORBmatcher.h) behind the same C harness that drives the unmodified reference ORBmatcher.cc
# (oracle/match_ref_harness.cc, oracle/matchshim); needs the reference tree for that header and DBoW2's FeatureVector —
# on the GPU box the prebuilt file shipped by gpurun is used
matchdropin: $(LIB)
	@if [ -f $(REF)/include/ORBmatcher.h ]; then \
	  mkdir -p tests/cpp/_build && \
	  g++ -O2 -std=c++14 -fPIC -shared -w -ffp-contract=off -Wl,-Bsymbolic -include oracle/matchshim/slam_types.h -Ioracle/matchshim \
	      -I$(REF)/include -I$(REF)/Thirdparty/DBoW2/DBoW2 -Iinclude -o $(DROPIN_M) oracle/match_ref_harness.cc \
	      $(PKG)/dropin/ORBmatcher.cc $(REF)/Thirdparty/DBoW2/DBoW2/FeatureVector.cpp \
	      -L$(PKG)/lib -leaof_orb -Wl,-rpath,'$$ORIGIN/../../../$(PKG)/lib' ; \
	else echo "reference tree $(REF) not present: keeping prebuilt $(DROPIN_M) (if any)"; fi

# the drop-in ORBVocabulary ($(PKG)/dropin/ORBVocabulary.h, a subclass of the reference's vendored DBoW2 vocabulary whose
# batch transform runs on the GPU) behind the harness that drives the unmodified DBoW2 (oracle/voc_ref_harness.cc)
DBOW := $(REF)/Thirdparty/DBoW2
vocdropin: $(LIB)
	@if [ -f $(DBOW)/DBoW2/TemplatedVocabulary.h ]; then \
	  mkdir -p tests/cpp/_build && \
	  /usr/bin/g++ -O2 -std=c++14 -fPIC -shared -w -Wl,-Bsymbolic -DEAOF_VOC_DROPIN -I$(PKG)/dropin -Ioracle/matchshim \
	      -I$(DBOW)/DBoW2 -I$(DBOW) -Iinclude -o $(DROPIN_V) oracle/voc_ref_harness.cc $(DBOW)/DBoW2/FORB.cpp \
	      $(DBOW)/DBoW2/BowVector.cpp $(DBOW)/DBoW2/FeatureVector.cpp $(DBOW)/DBoW2/ScoringObject.cpp \
	      $(DBOW)/DUtils/Random.cpp $(DBOW)/DUtils/Timestamp.cpp \
	      -L$(PKG)/lib -leaof_orb -Wl,-rpath,'$$ORIGIN/../../../$(PKG)/lib' ; \
	else echo "reference tree $(REF) not present: keeping prebuilt $(DROPIN_V) (if any)"; fi

oracle:
	$(MAKE) -C oracle
	$(MAKE) -C oracle ref

# compute-sanitizer over the GPU parity tests (run on a GPU box; logs of the last run: profiles/r01_sanitizer_*.log)
sanitize: all
	compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q
	compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "stages or voc or stereo or distinctive"

clean:
	rm -f $(LIB) build/ptxas.log $(DROPIN_T) $(DROPIN_M) $(DROPIN_V)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean matchdropin vocdropin sanitize
