// TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/ee_ref from the UNMODIFIED reference source.
//
// The reference translation unit is #included from where it lies under /root/reference
// (path passed as -DREF_EE_CPP=...), with its main() renamed.  This driver then calls the
// reference's own functions in the same order and with the same arguments as
// AmpliSolveErrorEstimation.cpp:426-454, skipping only generateReferenceBases
// (AmpliSolveErrorEstimation.cpp:578-670: one `samtools faidx` fork per panel position; no
// samtools and no hg19 FASTA exist in this image).  The two files that step would have
// produced (<seed>_panelReferenceBases.txt, <seed>_ampliconDuplicatedPositions.txt) are
// supplied by the caller instead.
//
// usage: ee_ref <panel.bed> <refbases.txt> <dups.txt> <germline_dir> <C_value> <cutoff> <out_dir> <list_file>
//        ee_ref default <panel.bed> <refbases.txt> <dups.txt> <default_error> <out_dir>
#define main ampli_reference_main_ee
#include REF_EE_CPP
#undef main

#include <chrono>

typedef std::chrono::steady_clock::time_point tp_t;
static double sec(tp_t a, tp_t b) { return std::chrono::duration<double>(b - a).count(); }

int main(int argc, char** argv) {
    if (argc == 7 && strcmp(argv[1], "default") == 0) {
        // AmpliSolveErrorEstimation.cpp:472-506 (germline_dir=not_available branch)
        float default_error = atof(argv[5]);
        if (default_error <= 0) default_error = 0.01;  // EE:349-363
        storeReference(argv[3], ReferenceBase_Hash);
        storeDuplicates(argv[4], DuplicatePosition_Hash);
        generateFinalOutput_default(0.0f, argv[2], ReferenceBase_Hash, DuplicatePosition_Hash, argv[6], default_error);
        return 0;
    }
    if (argc != 9) {
        fprintf(stderr, "usage: ee_ref panel refbases dups germline_dir C cutoff out_dir list_file\n");
        return 2;
    }
    char* panel = argv[1];
    float C_value_float = atof(argv[5]);          // EE:329
    int coverage_cutoff_int = atoi(argv[6]);      // EE:381
    if (C_value_float <= 0) C_value_float = 0.002;        // EE:372-376
    if (coverage_cutoff_int <= 0) coverage_cutoff_int = 100;  // EE:383-387
    tp_t t0 = std::chrono::steady_clock::now();
    storeReference(argv[2], ReferenceBase_Hash);
    storeDuplicates(argv[3], DuplicatePosition_Hash);
    generateCountList(argv[4], argv[8]);
    storeCountList(argv[8], argv[4], GermlineCountFileList_Hash);
    tp_t t1 = std::chrono::steady_clock::now();
    storeGermlineStatistics(GermlineCountFileList_Hash, GermlineValues_Hash_forThresholds, Germline_Max_Hash,
                            coverage_cutoff_int);
    tp_t t2 = std::chrono::steady_clock::now();
    estimateThresholds(C_value_float, ReferenceBase_Hash, GermlineValues_Hash_forThresholds,
                       Thresholds_Hash_Analytic, Count_Hash, Ratio_Hash, coverage_cutoff_int);
    tp_t t3 = std::chrono::steady_clock::now();
    generateFinalOutput(C_value_float, panel, ReferenceBase_Hash, DuplicatePosition_Hash, Thresholds_Hash_Analytic,
                        Germline_Max_Hash, argv[7], Count_Hash, Ratio_Hash);
    tp_t t4 = std::chrono::steady_clock::now();
    // machine-readable timing line used by bench.py's reference arm
    fprintf(stderr, "EE_REF_TIMING setup=%.6f parse=%.6f estimate=%.6f write=%.6f records=%zu\n", sec(t0, t1),
            sec(t1, t2), sec(t2, t3), sec(t3, t4), GermlineValues_Hash_forThresholds.size() / 4);
    return 0;
}
