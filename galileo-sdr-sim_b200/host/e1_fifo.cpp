/* e1_fifo.cpp -- see e1_fifo.h.  Restated from the behaviour of src/fifo.cpp and the producer block in
 * src/galileo-sdr.cpp:581-596 (head/tail arithmetic in samples, two-part copy at the wrap). */
#include "e1_fifo.h"

#include <string.h>

#include <condition_variable>
#include <mutex>
#include <new>

struct e1_fifo {
    int16_t *buf = nullptr;
    bool owned = false;
    size_t cap = 0;        /* samples; one slot is NOT sacrificed: `count` disambiguates full from empty */
    size_t head = 0, tail = 0, count = 0;
    bool finished = false;
    std::mutex m;
    std::condition_variable can_write, can_read;
};

extern "C" {

e1_fifo *e1_fifo_create(size_t capacity_samples, int16_t *storage)
{
    if (capacity_samples == 0)
        return nullptr;
    e1_fifo *f = new (std::nothrow) e1_fifo();
    if (!f)
        return nullptr;
    f->cap = capacity_samples;
    if (storage) {
        f->buf = storage;
    } else {
        f->buf = new (std::nothrow) int16_t[capacity_samples * 2];
        f->owned = true;
        if (!f->buf) {
            delete f;
            return nullptr;
        }
    }
    return f;
}

void e1_fifo_destroy(e1_fifo *f)
{
    if (!f)
        return;
    if (f->owned)
        delete[] f->buf;
    delete f;
}

size_t e1_fifo_sample_length(e1_fifo *f)
{
    std::lock_guard<std::mutex> g(f->m);
    return f->count;
}

static size_t read_locked(e1_fifo *f, int16_t *buffer, size_t samples)
{
    if (samples > f->count)
        samples = f->count;
    const size_t first = samples < f->cap - f->tail ? samples : f->cap - f->tail;
    memcpy(buffer, f->buf + f->tail * 2, first * 2 * sizeof(int16_t));
    memcpy(buffer + first * 2, f->buf, (samples - first) * 2 * sizeof(int16_t));
    f->tail = (f->tail + samples) % f->cap;
    f->count -= samples;
    return samples;
}

size_t e1_fifo_read(e1_fifo *f, int16_t *buffer, size_t samples)
{
    size_t n;
    {
        std::lock_guard<std::mutex> g(f->m);
        n = read_locked(f, buffer, samples);
    }
    if (n)
        f->can_write.notify_all();
    return n;
}

size_t e1_fifo_read_wait(e1_fifo *f, int16_t *buffer, size_t samples)
{
    size_t n;
    {
        std::unique_lock<std::mutex> g(f->m);
        const size_t want = samples < f->cap ? samples : f->cap;
        f->can_read.wait(g, [&] { return f->count >= want || f->finished; });
        n = read_locked(f, buffer, samples);
    }
    if (n)
        f->can_write.notify_all();
    return n;
}

int e1_fifo_write_ready(e1_fifo *f, size_t block_samples)
{
    std::lock_guard<std::mutex> g(f->m);
    return f->cap - f->count >= block_samples ? 1 : 0;
}

int e1_fifo_write(e1_fifo *f, const int16_t *iq, size_t samples)
{
    if (samples > f->cap)
        return -1;
    {
        std::unique_lock<std::mutex> g(f->m);
        f->can_write.wait(g, [&] { return f->cap - f->count >= samples || f->finished; });
        if (f->finished)
            return -1;
        const size_t first = samples < f->cap - f->head ? samples : f->cap - f->head;
        memcpy(f->buf + f->head * 2, iq, first * 2 * sizeof(int16_t));
        memcpy(f->buf, iq + first * 2, (samples - first) * 2 * sizeof(int16_t));
        f->head = (f->head + samples) % f->cap;
        f->count += samples;
    }
    f->can_read.notify_all();
    return 0;
}

void e1_fifo_finish(e1_fifo *f)
{
    {
        std::lock_guard<std::mutex> g(f->m);
        f->finished = true;
    }
    f->can_read.notify_all();
    f->can_write.notify_all();
}

int e1_fifo_finished(e1_fifo *f)
{
    std::lock_guard<std::mutex> g(f->m);
    return f->finished ? 1 : 0;
}

} /* extern "C" */
