// scheduler.cpp -- host-side request queues, KV page manager and the step-wise scheduler entry points
// of the reference API (include/item_storage.h, include/paged_item_storage.h).  These exist for code
// that drives the model one forward() at a time; the start_paged_* engines in engine.cpp do the same
// bookkeeping on the device instead.  Semantics follow the reference call for call:
//   process_decoder_result                    src/item_storage.cpp:97-139
//   non-paged insert_new_items                src/item_storage.cpp:141-180
//   paged insert_new_items                    src/paged_item_storage.cpp:62-122  (incl. quirk Q1 unless
//                                             mli::set_fix_stale_lengths(true))
//   allocate_or_free_memory_blocks_if_needed  src/paged_item_storage.cpp:14-60
//   MemoryBlockManager / PagedAttentionsManager  src/paged_item_storage.cpp:125-203
#include <algorithm>
#include <cassert>
#include <iterator>

#include "mli/compat.hpp"

// ---- Storage / ItemStorage / ProcessingStorage -----------------------------------------------------
std::vector<IdTokensPair> Storage::pop_pairs(int size) {
    std::vector<IdTokensPair> out;
    while (size-- > 0 && !data_.empty()) {
        out.push_back(std::move(data_.front()));
        data_.pop_front();
    }
    return out;
}
void Storage::add(IdTokensPair&& p) { data_.push_back(std::move(p)); }
void Storage::add_to_front(IdTokensPair&& p) { data_.push_front(std::move(p)); }
int Storage::size() const { return static_cast<int>(data_.size()); }
int Storage::head_length() const { return static_cast<int>(data_.front().second.size()); }
const IdTokensPair& Storage::get_top() const { return data_.front(); }
const std::list<IdTokensPair>& Storage::get_data() const { return data_; }

std::vector<IdTokensPair> ItemStorage::pop_finished_items(int n) { return finished_items_.pop_pairs(n); }
std::vector<IdTokensPair> ItemStorage::pop_new_items(int n) { return new_items_.pop_pairs(n); }
const IdTokensPair& ItemStorage::get_top() const { return new_items_.get_top(); }
void ItemStorage::add_finished_item(IdTokensPair&& p) { finished_items_.add(std::move(p)); }
void ItemStorage::add_new_item(IdTokensPair&& p) { new_items_.add(std::move(p)); }
void ItemStorage::add_new_item_to_head(IdTokensPair&& p) { new_items_.add_to_front(std::move(p)); }
int ItemStorage::finish_count() const { return finished_items_.size(); }
int ItemStorage::new_count() const { return new_items_.size(); }
int ItemStorage::head_length() const { return new_items_.head_length(); }
const std::list<IdTokensPair>& ItemStorage::get_finished_items() const { return finished_items_.get_data(); }

void ProcessingStorage::put(int row, IdTokensPair&& p) { batch_id_to_token_pairs_[row] = std::move(p); }
void ProcessingStorage::remove(int row) { batch_id_to_token_pairs_.erase(row); }
bool ProcessingStorage::batch_id_processing(int row) { return batch_id_to_token_pairs_.count(row) != 0; }
IdTokensPair& ProcessingStorage::get_token(int row) { return batch_id_to_token_pairs_[row]; }
int ProcessingStorage::size() const { return static_cast<int>(batch_id_to_token_pairs_.size()); }
void ProcessingStorage::move_to_finished(int row, ItemStorage& items) {
    auto it = batch_id_to_token_pairs_.find(row);
    items.add_finished_item(std::move(it->second));
    batch_id_to_token_pairs_.erase(it);
}
void ProcessingStorage::move_to_new(int row, ItemStorage& items) {
    auto it = batch_id_to_token_pairs_.find(row);
    items.add_new_item_to_head(std::move(it->second));
    batch_id_to_token_pairs_.erase(it);
}

void append_token_to_id_string_pair(IdTokensPair& p, int tok) { p.second.push_back(tok); }

bool is_done(ItemStorage& items, ProcessingStorage& processing) {
    return processing.size() + items.new_count() == 0;
}

// ---- process_decoder_result ----------------------------------------------------------------------------
std::vector<int> process_decoder_result(const TensorInt& decoder_result_device,
                                        TensorInt& decoder_result_host, ItemStorage& items,
                                        ProcessingStorage& processing, int n_sequence) {
    const auto& shp = decoder_result_host.shape();
    const int rows = static_cast<int>(shp[0]);
    const int rounds = shp.size() == 2 ? static_cast<int>(shp[1]) : 1;
    decoder_result_host.copy_from(decoder_result_device);  // the blocking D2H of the step
    const int* tok = decoder_result_host.data();
    std::vector<int> free_rows;
    int appended = 0;
    for (int r = 0; r < rows; ++r) {
        bool empty = false, finished = false;
        for (int j = 0; j < rounds && !empty && !finished; ++j) {
            const int t = tok[r * rounds + j];
            if (t == EMPTY_ROW_TOKEN_ID) {
                empty = true;
                continue;
            }
            IdTokensPair& req = processing.get_token(r);
            req.second.push_back(t);
            ++appended;
            finished = static_cast<int>(req.second.size()) >= n_sequence || t == EOF_TOKEN_ID;
        }
        if (empty || finished) free_rows.push_back(r);
        if (finished) processing.move_to_finished(r, items);
    }
    get_global_throughput_counter().add_record_if_recording(appended);
    return free_rows;
}

// ---- non-paged insert_new_items ---------------------------------------------------------------------------
int insert_new_items(const std::vector<int>& free_rows, TensorInt& inp_device, TensorInt& inp_host,
                     TensorInt& lengths_device, TensorInt& lengths_host, TensorInt& idx_device,
                     TensorInt& idx_host, ItemStorage& items, ProcessingStorage& processing) {
    if (free_rows.empty()) return 0;
    std::vector<IdTokensPair> fresh = items.pop_new_items(static_cast<int>(free_rows.size()));
    inp_host.copy_from(inp_device);
    lengths_host.copy_from(lengths_device);  // device-side lengths are authoritative here
    const int S = static_cast<int>(inp_host.shape()[1]);
    int* inp = inp_host.data();
    int* len = lengths_host.data();
    int* idx = idx_host.data();
    for (size_t i = 0; i < free_rows.size(); ++i) {
        const int row = free_rows[i];
        idx[i] = row;
        if (i >= fresh.size()) {
            len[row] = 0;
            continue;
        }
        assert(static_cast<int>(fresh[i].second.size()) + 1 <= S);
        len[row] = static_cast<int>(fresh[i].second.size());
        std::copy(fresh[i].second.begin(), fresh[i].second.end(), inp + static_cast<size_t>(row) * S);
        processing.put(row, std::move(fresh[i]));
    }
    inp_device.copy_from(inp_host);
    lengths_device.copy_from(lengths_host);
    idx_device.copy_from(idx_host);
    return static_cast<int>(fresh.size());
}

// ---- MemoryBlockManager ---------------------------------------------------------------------------------------
MemoryBlockManager::MemoryBlockManager(int n_blocks, size_t each_block_size)
    : block_memory_({static_cast<size_t>(n_blocks) * each_block_size}, DeviceType::DEVICE),
      n_blocks_(n_blocks), each_block_size_(each_block_size) {
    float* base = block_memory_.data();
    for (int i = 0; i < n_blocks; ++i) free_blocks_.push_back(base + static_cast<size_t>(i) * each_block_size);
}
int MemoryBlockManager::free_blocks_size() const { return static_cast<int>(free_blocks_.size()); }
void MemoryBlockManager::return_free_blocks(std::list<float*>&& blocks) {
    free_blocks_.splice(free_blocks_.end(), blocks);  // FIFO: returned pages go to the back
}
std::list<float*> MemoryBlockManager::pop_free_blocks(int size) {
    if (free_blocks_size() < size) throw std::runtime_error("No enough block memories to return");
    std::list<float*> out;
    auto last = free_blocks_.begin();
    std::advance(last, size);
    out.splice(out.end(), free_blocks_, free_blocks_.begin(), last);
    return out;
}

// ---- PagedAttentionsManager ---------------------------------------------------------------------------------------
PagedAttentionsManager::PagedAttentionsManager(size_t max_batches, size_t n_sequence, size_t)
    : page_table_host({max_batches, n_sequence / PAGE_BLOCK_SIZE}, DeviceType::HOST),
      page_table_device({max_batches, n_sequence / PAGE_BLOCK_SIZE}, DeviceType::DEVICE),
      width_(n_sequence / PAGE_BLOCK_SIZE), needs_sync_(false) {
    assert(n_sequence % PAGE_BLOCK_SIZE == 0);
}
std::list<BatchIdMemoryBlocksPair>& PagedAttentionsManager::get_used_block_list() { return used_blocks_; }
TensorFloatPoint& PagedAttentionsManager::get_page_table_device() { return page_table_device; }
void PagedAttentionsManager::maybe_flush_changes() {
    if (needs_sync_) page_table_device.copy_from(page_table_host);
    needs_sync_ = false;
}
void PagedAttentionsManager::set_block_pos(int row, int i_block, float* page) {
    page_table_host.data()[static_cast<size_t>(row) * width_ + i_block] = page;
    needs_sync_ = true;
}
void PagedAttentionsManager::add_batch_block_pair(BatchIdMemoryBlocksPair&& pair) {
    float** table_row = page_table_host.data() + static_cast<size_t>(pair.first) * width_;
    size_t i = 0;
    for (float* page : pair.second) table_row[i++] = page;
    used_blocks_.push_back(std::move(pair));
    needs_sync_ = true;
}

void allocate_memory_block(MemoryBlockManager& pool, PagedAttentionsManager& pages,
                           BatchIdMemoryBlocksPair& row) {
    float* page = pool.pop_free_blocks(1).front();
    row.second.push_front(page);
    // the table slot is the new page count - 1, regardless of where the list node went
    pages.set_block_pos(row.first, static_cast<int>(row.second.size()) - 1, page);
}

// ---- allocate_or_free_memory_blocks_if_needed ---------------------------------------------------------
void allocate_or_free_memory_blocks_if_needed(PagedAttentionsManager& pages, MemoryBlockManager& pool,
                                              ProcessingStorage& processing, ItemStorage& items,
                                              const std::vector<int>& free_rows, int rounds) {
    assert(rounds > 0 && rounds <= PAGE_BLOCK_SIZE);
    auto& used = pages.get_used_block_list();
    // retire: rows reported finished or empty give their pages back, in list order
    for (auto it = used.begin(); it != used.end();) {
        if (std::find(free_rows.begin(), free_rows.end(), it->first) != free_rows.end()) {
            pool.return_free_blocks(std::move(it->second));
            it = used.erase(it);
        } else {
            ++it;
        }
    }
    // grow by one page where the next `rounds` tokens would not fit; when the pool is dry the
    // TAIL of the list is pre-empted back to the front of the queue and the row is looked at again
    for (auto it = used.begin(); it != used.end();) {
        const size_t tokens = processing.get_token(it->first).second.size();
        if (tokens + rounds <= it->second.size() * PAGE_BLOCK_SIZE) {
            ++it;
        } else if (pool.free_blocks_size() > 0) {
            allocate_memory_block(pool, pages, *it);
        } else if (std::next(it) == used.end()) {
            processing.move_to_new(it->first, items);
            pool.return_free_blocks(std::move(it->second));
            it = used.erase(it);
        } else {
            BatchIdMemoryBlocksPair victim(std::move(used.back()));
            used.pop_back();
            processing.move_to_new(victim.first, items);
            pool.return_free_blocks(std::move(victim.second));
        }
    }
}

// ---- paged insert_new_items ------------------------------------------------------------------------------------------
std::vector<int> insert_new_items(TensorInt& inp_device, TensorInt& inp_host, TensorInt& lengths_device,
                                  TensorInt& lengths_host, TensorInt& idx_device, TensorInt& idx_host,
                                  ItemStorage& items, ProcessingStorage& processing,
                                  MemoryBlockManager& pool, PagedAttentionsManager& pages, int rounds) {
    assert(rounds > 0 && rounds <= PAGE_BLOCK_SIZE);
    const int B = static_cast<int>(inp_device.shape()[0]);
    const int S = static_cast<int>(inp_device.shape()[1]);
    std::vector<char> occupied(B, 0);
    for (const auto& row : pages.get_used_block_list()) occupied[row.first] = 1;
    // the reference never refreshes lengths_host from the device (quirk Q1); the corrected mode does
    if (mli::fix_stale_lengths()) lengths_host.copy_from(lengths_device);
    int* inp = inp_host.data();
    int* len = lengths_host.data();
    int* idx = idx_host.data();
    std::vector<int> admitted;
    bool touched = false;
    for (int row = 0; row < B; ++row) {
        if (occupied[row]) continue;
        touched = true;
        const bool fits = pool.free_blocks_size() >= DEFAULT_INIT_NUM_BLOCKS && items.new_count() > 0 &&
                          pool.free_blocks_size() >= ceil_div(items.head_length() + rounds, PAGE_BLOCK_SIZE);
        if (!fits) {
            len[row] = 0;
            continue;
        }
        IdTokensPair req = std::move(items.pop_new_items(1)[0]);
        const int n = static_cast<int>(req.second.size());
        assert(n + 1 <= S);
        len[row] = n;
        std::copy(req.second.begin(), req.second.end(), inp + static_cast<size_t>(row) * S);
        idx[admitted.size()] = row;
        const int n_pages = std::max(ceil_div(n + rounds, PAGE_BLOCK_SIZE), DEFAULT_INIT_NUM_BLOCKS);
        processing.put(row, std::move(req));
        pages.add_batch_block_pair(std::make_pair(row, pool.pop_free_blocks(n_pages)));
        admitted.push_back(row);
    }
    if (touched) {
        inp_device.copy_from(inp_host);
        lengths_device.copy_from(lengths_host);
        idx_device.copy_from(idx_host);
    }
    pages.maybe_flush_changes();
    return admitted;
}
