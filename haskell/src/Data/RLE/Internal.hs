{-# LANGUAGE DeriveGeneric #-}
{-# LANGUAGE TypeFamilies  #-}

-- |
-- Module      :  Data.RLE.Internal
-- Description :  drop-in replacement of text-compression's Data.RLE.Internal over the B200 kernels
--
-- Export list and types of the reference (src/Data/RLE/Internal.hs:43-50).  A run is rendered as the two
-- elements @Just (decimal count), symbol@ exactly as the reference does, including its treatment of 'Nothing'
-- (a 'Nothing' closes the current run WITHOUT resetting the count, a leading 'Nothing' is dropped, a trailing
-- one flushes twice: SURVEY.md 2.3 Q1-Q3).  Sequences whose items are single bytes -- everything the
-- ByteString / Text wrappers of "Data.RLE" produce -- run on the device (tc_rle_encode / tc_rle_decode); other
-- items are rank-compressed onto bytes when at most 256 distinct ones occur, else handled on the host.
-- NOT COMPILED: no GHC exists in the build image (see "Data.TextCompression.B200").
module Data.RLE.Internal ( Pack(pck, unpck, Itm, one),
                           -- * Base RLE types
                           RLE(..),
                           -- * To RLE functions
                           seqToRLE,
                           -- * From RLE functions
                           seqFromRLE,
                           -- * Used by "Data.MTF.Internal" and "Data.FMIndex.Internal" (not part of the reference's list)
                           symbolBytes
                         ) where

import           Data.ByteString              (ByteString)
import qualified Data.ByteString              as BS
import qualified Data.ByteString.Char8        as BSC8
import           Data.Foldable                (foldl', toList)
import qualified Data.Map.Strict              as M
import           Data.Maybe                   (catMaybes, fromJust, isJust, isNothing)
import           Data.Sequence                (Seq (..), (|>))
import qualified Data.Sequence                as DS
import           Data.Text                    (Text)
import qualified Data.Text                    as DText
import qualified Data.Text.Encoding           as DTE
import qualified Data.TextCompression.B200    as B200
import           Data.Word                    (Word8)
import           GHC.Generics                 (Generic)

-- | Items a run-length stream can carry.  'pck', 'unpck', 'Itm' and 'one' are the reference's interface;
-- 'fromString' / 'toString' render the counts; 'sortKey' (not exported, defaulted) gives the instances of this
-- module a total order so that their items can be rank-compressed onto bytes.
class (Monoid b, Eq b) => Pack b where
  pck :: [Itm b] -> b
  unpck :: b -> [Itm b]
  type Itm b
  one :: Itm b -> b
  fromString :: String -> b
  toString :: b -> String
  sortKey :: b -> Maybe ByteString
  sortKey _ = Nothing

instance Pack ByteString where
  pck = BS.pack
  unpck = BS.unpack
  type Itm ByteString = Word8
  one = BS.singleton
  fromString = BSC8.pack
  toString = BSC8.unpack
  sortKey = Just

instance Pack Text where
  pck = DText.pack
  unpck = DText.unpack
  type Itm Text = Char
  one = DText.singleton
  fromString = pck
  toString = unpck
  sortKey = Just . DTE.encodeUtf8

newtype RLE b = RLE (Seq (Maybe b))
  deriving (Eq,Ord,Show,Read,Generic)

-- | The sequence over byte symbols plus the way back, when its items can be told apart by bytes: items that ARE
-- single bytes map to themselves (so the alphabet order of "Data.MTF" is kept), any other item set of at most
-- 256 distinct values is ranked in the order of its keys.
symbolBytes :: Pack b => Seq (Maybe b) -> Maybe (Seq (Maybe Word8), Word8 -> b)
symbolBytes xs = do
  keys <- traverse (traverse sortKey) xs
  let items    = catMaybes (toList keys)
      distinct = M.fromList (zip items (catMaybes (toList xs)))        -- key -> an item with that key
  if all ((== 1) . BS.length) items
    then Just (fmap (fmap BS.head) keys, \w -> distinct M.! BS.singleton w)
    else if M.size distinct > 256 then Nothing else
      let ranks = M.fromDistinctAscList (zip (M.keys distinct) [0 ..]) :: M.Map ByteString Word8
          back  = DS.fromList (M.elems distinct)
      in Just (fmap (fmap (ranks M.!)) keys, DS.index back . fromIntegral)

-- | Runs as (count, symbol) pairs, with the reference's state machine (host version, any 'Eq' item).
hostRuns :: Eq b => Seq (Maybe b) -> Seq (Int, Maybe b)
hostRuns DS.Empty      = DS.empty
hostRuns (x :<| rest)  = let (out, c, item) = foldl' step (DS.empty, 1, x) rest in out |> (c, item)
  where
    step (out, c, item) y
      | isNothing y    = (out |> (c, item) |> (1, Nothing), c, Nothing)   -- the count survives a Nothing
      | isNothing item = (out, 1, y)                                      -- nothing is emitted for the Nothing itself
      | y == item      = (out, c + 1, item)
      | otherwise      = (out |> (c, item), 1, y)

seqToRLE :: Pack b => Seq (Maybe b) -> Seq (Maybe b)
seqToRLE DS.Empty = DS.empty
seqToRLE xs       = foldMap render runs
  where
    runs = case symbolBytes xs of
             Just (ws, back) -> fmap (fmap (fmap back)) (B200.seqToRLEW8 ws)   -- flag / scan / compact on the GPU
             Nothing         -> hostRuns xs
    render (c, s) = DS.fromList [Just (fromString (show c)), s]

seqFromRLE :: Pack b => RLE b -> Seq (Maybe b)
seqFromRLE (RLE DS.Empty) = DS.empty
seqFromRLE (RLE ys)       =
  case symbolBytes (fmap snd runs) of
    Just (ws, back) | all (isJust . fst) runs ->
      fmap (fmap back) (B200.seqFromRLEW8 (DS.zip (fmap (read . toString . fromJust . fst) runs) ws))
    _ -> foldMap expand runs
  where
    runs = pairs ys
    pairs (a :<| b :<| more) = (a, b) :<| pairs more
    pairs _                  = DS.empty                        -- an odd trailing element is ignored
    expand (y1, y2)
      | isJust y1 && isNothing y2 = DS.singleton Nothing       -- one Nothing, whatever the count says
      | otherwise                 = DS.replicate (read (toString (fromJust y1))) y2
