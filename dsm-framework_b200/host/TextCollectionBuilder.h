// TextCollectionBuilder.h -- drop-in for the reference's TextCollectionBuilder
// (TextCollectionBuilder.h:41-73): same constructor arguments, InsertText
// overloads and InitTextCollection, same error behaviour (message on cerr and
// exit(1) for insert-after-init and empty texts, TextCollectionBuilder.cpp:67-71,
// 86-91).  Behind it the documents stream to the GPU through the C ABI of
// include/dsmfm.h instead of into incbwt's RLCSABuilder.
#ifndef DSMFM_HOST_TEXTCOLLECTIONBUILDER_H_
#define DSMFM_HOST_TEXTCOLLECTIONBUILDER_H_

#include "TextCollection.h"
#include <cstring>
#include <string>

// Default samplerate for suffix array samples
#define TEXTCOLLECTION_DEFAULT_SAMPLERATE 124

// Default input length (the reference sizes its 512 MiB batch buffer from it;
// here it is only a hint for the device text buffer)
#define TEXTCOLLECTION_DEFAULT_INPUT_LENGTH (5lu * 1024 * 1024 * 1024)

struct TCBuilderRep; // Pimpl

class TextCollectionBuilder
{
public:
    explicit TextCollectionBuilder(unsigned samplerate = TEXTCOLLECTION_DEFAULT_SAMPLERATE,
                                   ulong estimatedInputLength = TEXTCOLLECTION_DEFAULT_INPUT_LENGTH,
                                   TextCollection::IndexType type = TextCollection::TYPE_FMINDEX);
    ~TextCollectionBuilder();

    // Insert a zero-terminated text from alphabet [1,255].  The i'th insertion
    // gets document identifier i-1.  Not allowed after InitTextCollection().
    void InsertText(uchar const *);
    void InsertText(uchar const *, std::string const &);

    // Extension (no counterpart in the reference class): the record loop of the reference CLI -- rows, '>' headers,
    // normalize() and transform(), builder.cpp:60-104, 183-262 -- run on the GPU over `length` bytes of FASTA text.
    // Every record with a non-empty sequence becomes one document, exactly as if the CLI had called InsertText
    // for it.  final = false (streaming): only the bytes in front of the last header line are consumed
    // (report.consumed); present the rest again followed by more input.
    struct FastaReport
    {
        ulong consumed, records, documents, bases, invalidRecords, firstInvalidOffset, badHeaders;
    };
    void InsertFasta(uchar const *text, ulong length, bool final, FastaReport &report);

    // Page-locked host memory for InsertFasta / InsertText input (faster host-to-device copies).
    static void *AllocPinned(ulong bytes);
    static void FreePinned(void *);

    // Build the static index on the GPU.  The caller deletes the result.
    TextCollection *InitTextCollection(bool storePlainText = false, bool color = false, unsigned rotationLength = 0);

private:
    struct TCBuilderRep *p_;
    TextCollectionBuilder(TextCollectionBuilder const &);
    TextCollectionBuilder &operator=(TextCollectionBuilder const &);
};

#endif
